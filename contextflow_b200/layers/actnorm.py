"""ActNorm with data-dependent initialisation and per-sample context shift / log-scale (reference layers/actnorm.py:7-102).
Reference quirks kept on purpose: ldj = +sum(logs) with no H*W factor; the first batch initialises the
parameters whenever `initialized == 0`, also in eval mode."""
import torch
import torch.nn as nn

from .. import ops, training
from .activations import FlowActivationLayer
from .context import ContextPlan
from .flowlayer import PackCache, inference_only

__all__ = ['ActNorm', 'ActNormFC']


class ActNorm(FlowActivationLayer):
    stats_fn = None          # set by sharded.enable_sharded_actnorm_init: global-batch statistics over all ranks

    def __init__(self, data_size, context_net=None, contextflow=False):
        super().__init__()
        D, H, W = data_size if len(data_size) == 3 else (data_size[0], 1, 1)
        self.D, self.H, self.W = D, H, W
        self.NN_t = nn.Parameter(torch.zeros(D))
        self.NN_logs = nn.Parameter(torch.zeros(D))
        self.register_buffer('initialized', torch.tensor(0))
        self.context_net = context_net
        self.contextflow = contextflow
        if self.context_net:
            self.C = C = self.context_net.C
            self.CN = nn.Linear(C, 2 * D)
            if self.contextflow:
                self.NN_t.requires_grad_(False)
                self.NN_logs.requires_grad_(False)
                nn.init.zeros_(self.CN.weight)
                nn.init.zeros_(self.CN.bias)
        self._plan, self._packs = ContextPlan(), PackCache()
        self._init_seen = False
        self._register_load_state_dict_pre_hook(self._forget_init)

    def _forget_init(self, *args, **kwargs):
        self._init_seen = False

    def is_initialized(self):
        if not self._init_seen:                      # one device read until the flag flips, none afterwards
            self._init_seen = bool(self.initialized.item())
        return self._init_seen

    def initialize(self, x):
        """actnorm.py:28-35: t <- mean, logs <- log(unbiased std + 1e-8) over (B,H,W)."""
        with torch.no_grad():
            mean, logstd = (type(self).stats_fn or ops.actnorm_stats)(x)     # sharded first batch: contextflow_b200.sharded
            self.NN_t.data.copy_(mean)
            self.NN_logs.data.copy_(logstd)
            self.initialized.fill_(1)
        self._init_seen = True

    def context_affine(self, context):
        c, logp_c = self._plan.run(self.context_net, context)
        wt = self._packs.get('cn', [self.CN.weight], lambda: ops.pack_kmajor(self.CN.weight, 1))
        return ops.linear(c, wt, self.CN.bias.detach()), logp_c          # (B, 2D) 'b (p d)'

    def forward(self, x, context=None):
        if self.context_net and training.wants_grad(x, self.CN.weight, self.CN.bias):
            if self.contextflow:                                       # conventional: NN_t / NN_logs are unused (actnorm.py:53-54) and get no gradient
                inference_only(self.NN_t)
            if self.contextflow and not self.is_initialized():
                self.initialize(x)
            c, logp_c = training.encode(self, context)
            cm = training.LinearRowsFn.apply(c, self.CN.weight, self.CN.bias)
            return training.ActNormCtxFn.apply(x, cm, logp_c, self)
        HW = x.shape[2] * x.shape[3]
        if self.context_net:
            cm, logp_c = self.context_affine(context)
            if self.contextflow:
                if not self.is_initialized():
                    self.initialize(x)
                return ops.actnorm(x, self.NN_t.detach(), self.NN_logs.detach(), cm, logp_c, float(HW), mode=1)
            return ops.actnorm(x, None, None, cm, logp_c, float(HW), mode=2)
        if not self.is_initialized():
            self.initialize(x)
        if training.wants_grad(x, self.NN_t, self.NN_logs):
            return training.ActNormFn.apply(x, self.NN_t, self.NN_logs)
        return ops.actnorm(x, self.NN_t.detach(), self.NN_logs.detach())

    def reverse(self, z, context=None):
        """actnorm.py:62-79: x = z * exp(logs) + t.  The reference's context branch calls rearrange on the (c, logp_c) tuple its
        context_net returns (:65) and cannot run, so only the context-free form exists."""
        inference_only(self.NN_t); inference_only(z)
        if self.context_net:
            raise NotImplementedError('ActNorm.reverse with a context_net is not executable in the reference (actnorm.py:65); '
                                      'only context-free (generalist) layers invert')
        assert self.is_initialized()                                  # actnorm.py:74
        return ops.actnorm_inv(z, self.NN_t.detach(), self.NN_logs.detach())

    def logdet(self, x, context=None):
        return self.forward(x, context)[1]


class ActNormFC(ActNorm):
    def __init__(self, data_size):
        super().__init__(data_size)

    def forward(self, x, context=None):
        out, ldj = super().forward(x.reshape(-1, self.D, 1, 1), context)
        return out.view(-1, self.D), ldj

    def reverse(self, z, context=None):
        return super().reverse(z.reshape(-1, self.D, 1, 1), context).view(-1, self.D)

    def logdet(self, x, context=None):
        return self.forward(x, context)[1]
