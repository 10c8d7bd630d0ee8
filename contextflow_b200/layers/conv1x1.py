"""Invertible 1x1 convolution with optional per-sample context matrix (reference layers/conv1x1.py:9-96)."""
import torch
import torch.nn as nn

from .. import ops, training
from .context import ContextPlan
from .flowlayer import FlowLayer, PackCache, inference_only

__all__ = ['Conv1x1', 'FC']


class Conv1x1(FlowLayer):
    def __init__(self, data_size, context_net=None, contextflow=False):
        super().__init__()
        D, H, W = data_size if len(data_size) == 3 else (data_size[0], 1, 1)
        self.D, self.H, self.W = D, H, W
        self.NN = nn.Parameter(torch.Tensor(D, D))
        nn.init.orthogonal_(self.NN)
        self.context_net = context_net
        self.contextflow = contextflow
        if self.context_net:
            self.C = C = self.context_net[0].C if isinstance(self.context_net, list) else self.context_net.C
            self.CN = nn.Linear(C, D * D)
            nn.init.zeros_(self.CN.weight)
            nn.init.zeros_(self.CN.bias)
            if self.contextflow:
                self.NN.requires_grad_(False)
        self._plan, self._packs = ContextPlan(), PackCache()

    def logabsdet(self):
        """Device scalar log|det NN|, recomputed only when NN changes (the reference runs slogdet every forward)."""
        return self._packs.get('slogdet', [self.NN], lambda: ops.slogdet(self.NN.detach()))

    def context_matrix(self, context):
        c, logp_c = self._plan.run(self.context_net, context)
        wt = self._packs.get('cn', [self.CN.weight], lambda: ops.pack_kmajor(self.CN.weight, 1))
        return ops.linear(c, wt, self.CN.bias.detach()), logp_c          # (B, D*D) 'b (d1 d2)'

    def forward(self, x, context=None):
        if training.wants_grad(x, self.NN) and not self.context_net:
            return training.Conv1x1Fn.apply(x, self.NN, self)          # autograd through libcfpp kernels
        if self.context_net and training.wants_grad(x, self.CN.weight, self.CN.bias):
            if self.contextflow:                                       # specialist: gradients w.r.t. CN, the encoder and the input;
                inference_only(self.NN)                               # a conventional specialist never reads NN (conv1x1.py:46-49): it gets no gradient, as in the reference
            c, logp_c = training.encode(self, context)
            cmat = training.LinearRowsFn.apply(c, self.CN.weight, self.CN.bias)
            return training.Conv1x1CtxFn.apply(x, cmat, logp_c, self)
        inference_only(self.NN); inference_only(x)
        lad = self.logabsdet()
        if self.context_net:
            cm, logp_c = self.context_matrix(context)
            return ops.conv1x1(x, self.NN.detach(), lad, cm, logp_c, self.contextflow)
        return ops.conv1x1(x, self.NN.detach(), lad)

    def inverse_matrix(self, check=True):
        """torch.inverse(NN) (conv1x1.py:70), recomputed only when NN changes.  check=True (the sampling direction): a singular NN raises
        like torch.inverse does -- one device read per weight version; check=False (the training backward, every step a new version):
        no host synchronisation, a singular NN shows up as inf/nan gradients exactly as in torch."""
        inv, flag = self._packs.get('inverse', [self.NN], lambda: ops.mat_inverse(self.NN.detach()))
        if check and not getattr(self, '_inv_checked', None) == self.NN._version:
            if int(flag.item()):
                raise RuntimeError('Conv1x1.reverse: NN is singular')
            self._inv_checked = self.NN._version
        return inv

    def reverse(self, z, context=None):
        """conv1x1.py:59-72.  Only the context-free branch is executable in the reference: its context branch reads
        `w_ginv` before assigning it (:62) and multiplies by the nn.Linear module `self.CN`, so there is nothing to be on par with."""
        inference_only(self.NN); inference_only(z)
        if self.context_net:
            raise NotImplementedError('Conv1x1.reverse with a context_net is not executable in the reference (conv1x1.py:62); '
                                      'only context-free (generalist) layers invert')
        return ops.conv1x1(z, self.inverse_matrix(), self.logabsdet())[0]

    def logdet(self, input, context=None):
        return self.forward(input, context)[1]


class FC(Conv1x1):
    """Fully connected invertible layer on (B, D) vectors (inside context encoders)."""

    def __init__(self, data_size, context_net=None, contextflow=False):
        super().__init__(data_size, context_net=None, contextflow=False)

    def forward(self, x, context=None):
        out, ldj = super().forward(x.reshape(-1, self.D, 1, 1), context)
        return out.view(-1, self.D), ldj

    def reverse(self, z, context=None):
        return super().reverse(z.reshape(-1, self.D, 1, 1), context).view(-1, self.D)

    def logdet(self, x, context=None):
        return self.forward(x, context)[1]
