"""Fused execution plan of `FlowSequential.log_prob` (reference layers/flowsequential.py:18-30).

`FlowSequential.forward` keeps the reference's layer-by-layer contract (every module returns its own (z, ldj)).  `log_prob`
only needs the final (B, M) log-probabilities, so it may evaluate the same stack with fewer passes over HBM and fewer launches:

  * image prologue: Dequantization, Normalization x2, LogitTransform [, Augment] (model.py:97-100,121-123) -> one kernel;
  * context pre-pass: every layer's context encoder in one launch (EncoderBatch) and then every CN context network
    (coupling.py:37, actnorm.py:21, conv1x1.py:22) in one launch (cfpp_cn_batch); ActNorm's frozen NN_t / NN_logs are folded
    into its CN bias, so the per-sample (t | logs) rows come out of that launch ready to use;
  * Conv1x1 followed by ActNorm -> one kernel (the 1x1 conv with the ActNorm epilogue);
  * Coupling -> conditioner kernel + fused coupling kernel;
  * the `logdet += ldj` chain (flowsequential.py:23,27) -> one kernel after the last layer.

Draw order and shapes of the random numbers are those of the reference.  The plan is used only when it is
valid: inference (no autograd), every ActNorm initialised, recognised context encoders, no forward hooks on the layers; in every
other case `log_prob` runs the layer-by-layer forward, which is the same CUDA code one layer at a time.
"""
from __future__ import annotations

import os

import torch

from .. import ops
from .actnorm import ActNorm
from .augment import Augment
from .conv1x1 import Conv1x1
from .coupling import Coupling, TransCoupling
from .dequantize import Dequantization
from .normalize import Normalization
from .transforms import LogitTransform


def _has_hooks(m):
    return bool(m._forward_hooks) or bool(m._forward_pre_hooks)


class _Prologue:
    def __init__(self, deq, n0, n1, logit, aug):
        self.mods = [deq, n0, n1, logit] + ([aug] if aug is not None else [])
        self.deq, self.n0, self.n1, self.aug = deq, n0, n1, aug

    def run(self, x, ctx, terms):
        B, C = x.shape[0], x.shape[1]
        u = self.deq.dist.draw(B, device=x.device)                       # dequantize.py:15
        eps = self.aug.distribution.draw(B, device=x.device) if self.aug is not None else None    # augment.py:15
        s0, t0, _ = self.n0.host_constants(); s1, t1, _ = self.n1.host_constants()
        D = x.numel() / max(B, 1) / C
        const = self.n0.logdet_value(C, D) + self.n1.logdet_value(C, D)
        y, ldj = ops.prologue(x, u, eps, s0, t0, s1, t1, const)
        terms.append(ldj)
        return y


class _ConvAct:
    """Conv1x1 + ActNorm (conv1x1.py:28-57, actnorm.py:37-60) as one kernel."""

    def __init__(self, conv, an):
        self.mods = [conv, an]
        self.conv, self.an = conv, an
        self.c_conv = self.c_an = None          # (cn_out, logp_c) filled by the context pre-pass
        self._fused_key = self._fused = None

    def fused_cn(self):
        """True when the layer's own context network runs inside the conv kernel (cfpp_conv1x1_ctx_fwd): no (B, D, D) matrix in HBM."""
        conv = self.conv
        if not conv.context_net or not self.an.context_net or os.environ.get('CFPP_C1X1_CTX', '1') == '0':
            return False
        key = (conv.D, conv.H * conv.W, conv.C)
        if self._fused_key != key:
            self._fused_key, self._fused = key, ops.conv1x1_ctx_supported(8, *key)
        return self._fused

    def ctx_plans(self):
        return [m._plan for m in (self.conv, self.an) if m.context_net]

    def cn_jobs(self):
        conv, an = self.conv, self.an
        jobs = []
        if conv.context_net and not self.fused_cn():
            wt = conv._packs.get('cn', [conv.CN.weight], lambda: ops.pack_kmajor(conv.CN.weight, 1))
            jobs.append((conv._plan, [(wt, conv.CN.bias.detach())], conv.D, self, 'c_conv'))
        if an.context_net:
            wt = an._packs.get('cn', [an.CN.weight], lambda: ops.pack_kmajor(an.CN.weight, 1))
            if an.contextflow:                   # t = CN(c)[:D] + NN_t, logs = CN(c)[D:] + NN_logs   (actnorm.py:48-50)
                bias = an._packs.get('cn_bias_folded', [an.CN.bias, an.NN_t, an.NN_logs],
                                     lambda: an.CN.bias.detach() + torch.cat([an.NN_t.detach(), an.NN_logs.detach()]))
            else:
                bias = an.CN.bias.detach()
            jobs.append((an._plan, [(wt, bias)], 0, self, 'c_an'))
        return jobs

    def run(self, x, ctx, terms):
        conv, an = self.conv, self.an
        HW = x.shape[2] * x.shape[3]
        kw = {}
        if an.context_net:
            tl, lp = self.c_an
            kw = dict(an_t=tl, an_logs=None, an_logp_c=lp, an_logp_scale=float(HW))
        else:
            kw = dict(an_t=an.NN_t.detach(), an_logs=an.NN_logs.detach())
        if conv.context_net and self.fused_cn() and HW == conv.H * conv.W:
            e, lp = conv._plan.preset; conv._plan.preset = None
            wt, bt = conv._packs.get('cn_tril', [conv.CN.weight, conv.CN.bias], lambda: ops.pack_cn_tril(conv.CN.weight, conv.CN.bias, conv.D))
            z, ldj = ops.conv1x1_ctx(x, e, wt, bt, conv.NN.detach(), conv.logabsdet(), lp, conv.contextflow, **kw)
        elif conv.context_net:
            if self.c_conv is None:                  # fused plan chosen for another spatial size than the one that arrived: the two-kernel route
                e, lp = conv._plan.preset; conv._plan.preset = None
                wk = conv._packs.get('cn', [conv.CN.weight], lambda: ops.pack_kmajor(conv.CN.weight, 1))
                self.c_conv = (ops.linear(e, wk, conv.CN.bias.detach()), lp)
            cm, lp = self.c_conv
            z, ldj = ops.conv1x1(x, conv.NN.detach(), conv.logabsdet(), cm, lp, conv.contextflow, **kw)
        else:
            z, ldj = ops.conv1x1(x, conv.NN.detach(), conv.logabsdet(), **kw)
        self.c_conv = self.c_an = None
        terms.append(ldj)
        return z


class _Coup:
    def __init__(self, m):
        self.mods = [m]
        self.m = m
        self.c_cn = None

    def ctx_plans(self):
        return [self.m._plan] if self.m.context_net else []

    def cn_jobs(self):
        m = self.m
        if not m.context_net:
            return []
        lin = [m.CN[0], m.CN[2], m.CN[4]]
        packs = m._packs.get('cn', [l.weight for l in lin], lambda: [ops.pack_kmajor(l.weight, 1) for l in lin])
        return [(m._plan, [(p, l.bias.detach()) for p, l in zip(packs, lin)], 0, self, 'c_cn')]

    def run(self, x, ctx, terms):
        m = self.m
        if m.context_net:
            m._cn_preset = self.c_cn             # consumed by _CouplingBase._context_terms
            self.c_cn = None
        z, ldj = m(x, ctx)
        terms.append(ldj)
        return z


class _Generic:
    def __init__(self, m):
        self.mods = [m]
        self.m = m

    def run(self, x, ctx, terms):
        z, ldj = self.m(x, ctx)
        terms.append(ldj)
        return z


class FastLogProb:
    def __init__(self, seq):
        self.seq = seq
        mods = list(seq.sequence_modules)
        segs, i = [], 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (isinstance(m, Dequantization) and i + 3 < len(mods) and isinstance(mods[i + 1], Normalization)
                    and isinstance(mods[i + 2], Normalization) and isinstance(mods[i + 3], LogitTransform)
                    and hasattr(m.dist, 'draw')):
                aug = mods[i + 4] if i + 4 < len(mods) and isinstance(mods[i + 4], Augment) and mods[i + 4].split_dim == 1 else None
                segs.append(_Prologue(m, mods[i + 1], mods[i + 2], mods[i + 3], aug)); i += 5 if aug is not None else 4
            elif type(m) is Conv1x1 and type(nxt) is ActNorm and m.D == nxt.D:
                segs.append(_ConvAct(m, nxt)); i += 2
            elif type(m) in (Coupling, TransCoupling):
                segs.append(_Coup(m)); i += 1
            else:
                segs.append(_Generic(m)); i += 1
        self.segs = segs
        self._actnorms = [m for m in mods if isinstance(m, ActNorm)]
        self._jobs_key, self._jobs = None, None

    # ------------------------------------------------------------------------------------------------ validity
    def usable(self, x, context):
        if torch.is_grad_enabled() or not x.is_cuda or x.dim() != 4 or x.dtype != torch.float32:
            return False
        seq = self.seq
        if _has_hooks(seq) or any(_has_hooks(m) for m in seq.sequence_modules) or _has_hooks(seq.dist):
            return False
        if not all(a.is_initialized() for a in self._actnorms):
            return False
        groups = seq._encoder_groups() if context is not None else {}
        if len(groups) > 1:                      # one context pre-pass per forward (create_model stacks draw all layer noise up front)
            return False
        for s in self.segs:
            if isinstance(s, (_ConvAct, _Coup)):
                for plan in s.ctx_plans():
                    if context is None or context.dim() != 2:
                        return False
                    if not any(any(p is plan for p, _ in b.members) for b in groups.values()):
                        return False             # an encoder the batch launch does not cover: run layer by layer
            if isinstance(s, _Prologue):
                s0 = s.n0.scale.numel() == 1 and s.n0.translation.numel() == 1
                s1 = s.n1.scale.numel() == 1 and s.n1.translation.numel() == 1
                if not (s0 and s1):
                    return False
        return all(b.ready() for b in groups.values())

    # ------------------------------------------------------------------------------------------------ context pre-pass
    def _cn_descriptors(self):
        """[(ContextPlan, CnJob, owner segment, attribute)], rebuilt when a CN weight changes."""
        raw = [j for s in self.segs if isinstance(s, (_ConvAct, _Coup)) for j in s.cn_jobs()]
        key = tuple((wt.data_ptr(), None if b is None else (b.data_ptr(), b._version)) for _, layers, *_ in raw for wt, b in layers)
        if key != self._jobs_key:
            built = []
            for plan, layers, tril, seg, attr in raw:
                job, keep = ops.cn_job(layers, tril)
                built.append((plan, job, keep, seg, attr))
            self._jobs_key, self._jobs = key, built
        return self._jobs

    def _context_prepass(self, groups):
        jobs = self._cn_descriptors()
        if not jobs:
            return
        ins, lps = [], []
        for plan, *_ in jobs:
            c, lp = plan.preset; plan.preset = None
            ins.append(c); lps.append(lp)
        outs = ops.cn_batch([j for _, j, *_ in jobs], ins)
        for (plan, job, keep, seg, attr), o, lp in zip(jobs, outs, lps):
            setattr(seg, attr, (o, lp))

    # ------------------------------------------------------------------------------------------------ run
    def __call__(self, x, context=None):
        seq = self.seq
        B = x.shape[0]
        groups = seq._encoder_groups() if context is not None else {}
        terms, out, pos = [], x, 0
        for s in self.segs:
            for i in range(pos, pos + len(s.mods)):
                batch = groups.get(i)
                if batch is not None:             # the encoders of the layers from here on (noise drawn in layer order)
                    batch.run(context)
                    self._context_prepass(groups)
            out = s.run(out, context, terms)
            pos += len(s.mods)
        logprob = seq.dist.log_prob(out, context)
        return ops.ldj_sum(terms, B, seq.mixtures, x.device, last=logprob)
