"""Affine coupling layers (reference layers/coupling.py): Coupling (3-conv conditioner, :14-77), CouplingFC (:80-97),
TransCoupling (SimpleViT conditioner, :100-159).  Conditioner -> h in HBM -> fused coupling kernel (x,h -> z, ldj)."""
import torch
import torch.nn as nn

from .. import ops, training
from .context import ContextPlan
from .flowlayer import FlowLayer, PackCache, inference_only
from .simple_vit import SimpleViT

__all__ = ['Coupling', 'CouplingFC', 'TransCoupling', 'MaskedCoupling']


def _freeze(module):
    for p in module.parameters():
        p.requires_grad = False
    return module


def _context_mlp(C, H, O):
    return nn.Sequential(nn.Linear(C, H), nn.ReLU(), nn.Linear(H, H), nn.ReLU(), nn.Linear(H, O))


class _CouplingBase(FlowLayer):
    """Shared context plumbing: CN is Linear-ReLU-Linear-ReLU-Linear on the encoded context (coupling.py:37,121)."""

    def _context_terms(self, context):
        pre = getattr(self, '_cn_preset', None)                 # (CN(c), logp_c) from the fused plan's context pre-pass
        if pre is not None:
            self._cn_preset = None
            return pre
        c, logp_c = self._plan.run(self.context_net, context)
        lin = [self.CN[0], self.CN[2], self.CN[4]]
        packs = self._packs.get('cn', [l.weight for l in lin], lambda: [ops.pack_kmajor(l.weight, 1) for l in lin])
        h = ops.linear(c, packs[0], lin[0].bias.detach(), relu=True)
        h = ops.linear(h, packs[1], lin[1].bias.detach(), relu=True)
        return ops.linear(h, packs[2], lin[2].bias.detach()), logp_c

    def logdet(self, input, context=None):
        return self.forward(input, context)[1]


class Coupling(_CouplingBase):
    def __init__(self, data_channels, kernel_size=(1, 1), padding=(0, 0), context_net=None, contextflow=False):
        super().__init__()
        D, H, O = data_channels // 2, data_channels * 2, data_channels
        self.krn, self.pad = tuple(kernel_size), tuple(padding)
        self.context_net, self.contextflow = context_net, contextflow
        first_in = D + O if (context_net and not contextflow) else D          # conventional: concatenated context
        self.NN = nn.Sequential(nn.Conv2d(first_in, H, 1), nn.ReLU(),
                                nn.Conv2d(H, H, self.krn, padding=self.pad, padding_mode='reflect'), nn.ReLU(),
                                nn.Conv2d(H, O, 1))
        if self.context_net:
            if self.contextflow:
                _freeze(self.NN)
            self.C = self.context_net.C
            self.CN = _context_mlp(self.C, H, O)
        self._dims = (D, H, O)
        self._plan, self._packs = ContextPlan(), PackCache()
        if self.pad != (self.krn[0] // 2, self.krn[1] // 2):
            raise NotImplementedError('the fused conditioner assumes "same" reflect padding (model.py:114)')

    def _packed_nn(self):
        D, H, O = self._dims
        c1, c2, c3 = self.NN[0], self.NN[2], self.NN[4]

        def build():
            w1 = c1.weight.reshape(H, -1)
            return dict(main=(ops.pack_kmajor(w1[:, :D]), ops.pad_vec(c1.bias), ops.pack_kmajor(c2.weight.reshape(H, -1)),
                              ops.pad_vec(c2.bias), ops.pack_kmajor(c3.weight.reshape(O, -1)), ops.pad_vec(c3.bias)),
                        ctx_w=ops.pack_kmajor(w1[:, D:], 1) if w1.shape[1] > D else None,
                        tc=ops.conv_cond_tc_pack(c1.weight, c2.weight, c3.weight, D) if c1.weight.is_cuda else None)
        return self._packs.get('nn', [c1.weight, c1.bias, c2.weight, c2.bias, c3.weight, c3.bias], build)

    def _conditioner(self, x, pk, bias1_b=None):
        """h = NN(x0): tcgen05 kernel when the shape has a tensor-core plan, the FP32-FMA kernel otherwise (both CUDA, same result)."""
        D, H, O = self._dims
        Hh, Ww = x.shape[2], x.shape[3]
        if pk['tc'] is not None and ops.conv_cond_tc_mode() != 'fma':
            b1, b2, b3 = pk['main'][1], pk['main'][3], pk['main'][5]
            h = ops.conv_cond_tc(x, D, pk['tc'], b1, b2, b3, H, Hh, Ww, self.krn[0], self.krn[1], O, bias1_b=bias1_b)
            if h is not None:
                return h
        return ops.conv_cond(x, D, pk['main'], Hh, Ww, self.krn[0], self.krn[1], O, bias1_b=bias1_b)

    def _fused(self, x, pk, **kw):
        """Conditioner + coupling transform as one tensor-core kernel (h never reaches HBM); None when the shape has no fused plan."""
        # measured (B200, cfg2): round 1 the fused form lost (3.75 ms / step against 2.98 + 0.46 ms at B = 8192); with x1 prefetched ahead of
        # the stage-3 wait and the per-sample output bias staged in shared memory it ties at B = 8192 (3.20 vs 2.78 + 0.47 ms) and wins at
        # small batches, where every launch's fixed latency counts (B = 256: 0.682 vs 0.706 ms / step) -- 'auto' takes it below
        # ops.CONV_COND_FUSED_MAX_BATCH samples; CFPP_CONV_COND=fused / split force either route
        mode = ops.conv_cond_tc_mode()
        if pk['tc'] is None or not (mode == 'fused' or (mode == 'auto' and x.shape[0] < ops.CONV_COND_FUSED_MAX_BATCH)):
            return None
        b1, b2, b3 = pk['main'][1], pk['main'][3], pk['main'][5]
        return ops.conv_cond_tc_coupling(x, pk['tc'], b1, b2, b3, self._dims[1], self.krn[0], self.krn[1], **kw)

    def forward(self, x, context=None):
        if not self.context_net and training.wants_grad(x, *self.NN.parameters()):
            c1, c2, c3 = self.NN[0], self.NN[2], self.NN[4]
            return training.CouplingConvFn.apply(x, c1.weight, c1.bias, c2.weight, c2.bias, c3.weight, c3.bias)
        if self.context_net and training.wants_grad(x, *self.CN.parameters(), *self.NN.parameters()):
            if max(self._dims) > 320:
                raise NotImplementedError('CN wider than 320 features has no training kernel yet')
            c, logp_c = training.encode(self, context)
            lin = [self.CN[0], self.CN[2], self.CN[4]]
            cn = training.Mlp3RowsFn.apply(c, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias)
            c1, c2, c3 = self.NN[0], self.NN[2], self.NN[4]
            fn = training.CouplingCtxConvFn if self.contextflow else training.CouplingConcatConvFn     # additive (coupling.py:45) / concatenated (:47)
            return fn.apply(x, cn, logp_c, c1.weight, c1.bias, c2.weight, c2.bias, c3.weight, c3.bias)
        inference_only(x)
        D, H, O = self._dims
        Hh, Ww = x.shape[2], x.shape[3]
        pk = self._packed_nn()
        if not self.context_net:
            return self._fused(x, pk) or ops.coupling(x, self._conditioner(x, pk))
        cn, logp_c = self._context_terms(context)
        if self.contextflow:                                      # additive: h = NN(x0) + CN(c)   (coupling.py:45)
            return (self._fused(x, pk, add=cn, logp_c=logp_c, logp_scale=float(Hh * Ww))
                    or ops.coupling(x, self._conditioner(x, pk), add=cn, logp_c=logp_c, logp_scale=float(Hh * Ww)))
        # conventional: NN(cat(x0, CN(c) broadcast)) == first conv with the per-sample bias b1 + W1[:, D:] CN(c)  (coupling.py:47)
        bias1 = ops.linear(cn, pk['ctx_w'], self.NN[0].bias.detach())
        return (self._fused(x, pk, bias1_b=bias1, logp_c=logp_c, logp_scale=float(Hh * Ww))
                or ops.coupling(x, self._conditioner(x, pk, bias1_b=bias1), logp_c=logp_c, logp_scale=float(Hh * Ww)))


    def reverse(self, z, context=None):
        """coupling.py:68-73: the conditioner sees z0 = x0, so h is the forward pass's h; x1 = (z1 - t) / s."""
        inference_only(z)
        pk = self._packed_nn()
        if not self.context_net:
            return ops.coupling_inv(z, self._conditioner(z, pk))
        cn, _ = self._context_terms(context)
        if self.contextflow:
            return ops.coupling_inv(z, self._conditioner(z, pk), add=cn)
        bias1 = ops.linear(cn, pk['ctx_w'], self.NN[0].bias.detach())
        return ops.coupling_inv(z, self._conditioner(z, pk, bias1_b=bias1))


class CouplingFC(Coupling):
    def __init__(self, data_channels, kernel_size=(1, 1), padding=(0, 0), context_net=None, contextflow=False):
        super().__init__(data_channels, kernel_size=(1, 1), padding=(0, 0), context_net=None, contextflow=False)
        self.D = data_channels

    def forward(self, x, context=None):
        out, ldj = super().forward(x.reshape(-1, self.D, 1, 1), context)
        return out.view(-1, self.D), ldj

    def reverse(self, z, context=None):
        return super().reverse(z.reshape(-1, self.D, 1, 1), context).view(-1, self.D)

    def logdet(self, x, context=None):
        return self.forward(x, context)[1]


class TransCoupling(_CouplingBase):
    def __init__(self, in_sz, p_sz, context_net=None, contextflow=False):
        super().__init__()
        D, H, O = in_sz[0] // 2, in_sz[0] * 2, in_sz[0]
        T = O * p_sz[0] * p_sz[1]
        self.context_net, self.contextflow = context_net, contextflow
        vit = dict(image_size=(in_sz[1], in_sz[2]), patch_size=tuple(p_sz), dim=T, depth=6, heads=1, mlp_dim=T)
        if self.context_net and not self.contextflow:
            self.NN = SimpleViT(channels=D + O, **vit)              # bare module: keys NN.* (App. C-6)
        else:
            self.NN = nn.Sequential(SimpleViT(channels=D, **vit))   # keys NN.0.*
            if self.context_net:
                _freeze(self.NN)
        if self.context_net:
            self.C = self.context_net.C
            self.CN = _context_mlp(self.C, H, O)
        self._dims = (D, H, O)
        self._plan, self._packs = ContextPlan(), PackCache()

    def forward(self, x, context=None):
        vit = self.NN if isinstance(self.NN, SimpleViT) else self.NN[0]
        if not self.context_net and training.wants_grad(x, *vit.parameters()):
            return training.CouplingVitFn.apply(x, None, None, vit, *vit._sources())   # autograd through libcfpp kernels
        if self.context_net and training.wants_grad(x, *self.CN.parameters(), *vit.parameters()):
            if max(self._dims) > 320:
                raise NotImplementedError('CN wider than 320 features has no training kernel yet')
            c, logp_c = training.encode(self, context)
            lin = [self.CN[0], self.CN[2], self.CN[4]]
            cn = training.Mlp3RowsFn.apply(c, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias)
            return training.CouplingVitFn.apply(x, cn, logp_c, vit, *vit._sources())       # additive (--contextflow) or concatenated by the ViT's width
        inference_only(x)
        if not self.context_net:
            return ops.coupling(x, vit(x))
        cn, logp_c = self._context_terms(context)                   # note: logp_c is NOT scaled by H*W here (coupling.py:126)
        if self.contextflow:
            return ops.coupling(x, vit(x), add=cn, logp_c=logp_c, logp_scale=1.0)
        return ops.coupling(x, vit(x, extra=cn), logp_c=logp_c, logp_scale=1.0)

    def reverse(self, z, context=None):
        """coupling.py:150-155."""
        inference_only(z)
        vit = self.NN if isinstance(self.NN, SimpleViT) else self.NN[0]
        if not self.context_net:
            return ops.coupling_inv(z, vit(z))
        cn, _ = self._context_terms(context)
        if self.contextflow:
            return ops.coupling_inv(z, vit(z), add=cn)
        return ops.coupling_inv(z, vit(z, extra=cn))


class MaskedCoupling(FlowLayer):
    """`--coupling maf` (reference layers/ar.py:15-69): h = MaskedResidualBlock2d(x); t, r = halves of h; z = x * exp(2 tanh(r/2)) + t over
    all channels; ldj = sum log_s.  --contextflow specialist (ar.py:23-28,39-42): NN frozen, h += CN(c) with CN the masked residual linear
    block (which the reference can only evaluate when the encoder width is 1 or 2 * channels), ldj += H W logp_c.  The conventional
    specialist is not executable in the reference (its 3D-channel concatenation meets a 2D-channel conv1, ar.py:26,44) and raises here too."""

    def __init__(self, data_channels, kernel_size=(1, 1), padding=(0, 0), context_net=None, contextflow=False, mask_type='B'):
        super().__init__()
        from .autoregressive import MaskedResidualBlock2d, MaskedResidualBlockLinear
        D = data_channels
        self.context_net, self.contextflow = context_net, contextflow
        self.NN = MaskedResidualBlock2d(D, D, kernel_size=kernel_size, padding=padding, D=D, mask_type=mask_type)
        if self.context_net:
            if not self.contextflow:
                self.NN = MaskedResidualBlock2d(2 * D, D, kernel_size=kernel_size, padding=padding, D=D, mask_type=mask_type)
            else:
                _freeze(self.NN)
            self.C = self.context_net.C
            self.CN = MaskedResidualBlockLinear(self.C, D, D)
            self._plan = ContextPlan()

    def forward(self, x, context=None):
        nn_ = self.NN
        if self.context_net:
            if not self.contextflow:
                raise RuntimeError('MaskedCoupling without --contextflow concatenates 2D context channels to the D data channels and feeds a '
                                   '2D-channel masked conv (ar.py:26,44): the reference raises a channel mismatch here as well')
            if training.wants_grad(x, *self.CN.parameters()):
                raise NotImplementedError('training a MaskedCoupling specialist has no backward kernel (only the forward is built)')
            inference_only(x)
            c, logp_c = self._plan.run(self.context_net, context)
            return ops.maf_coupling(x, nn_(x, identity=False), add=self.CN(c), logp_c=logp_c, logp_scale=float(x.shape[2] * x.shape[3]))
        if training.wants_grad(x, *nn_.parameters()):
            for c in (nn_.conv1, nn_.conv2, nn_.conv3):
                c.masked_weight()                                                                  # mask in place first (masked_conv_2d.py:21-23)
            ws = [nn_.conv1.weight, nn_.conv2.weight, nn_.conv3.weight]
            return training.MaskedCouplingFn.apply(x, ws[0], nn_.conv1.bias, ws[1], nn_.conv2.bias, ws[2], nn_.conv3.bias)
        return ops.maf_coupling(x, nn_(x, identity=False))

    def reverse(self, z, context=None):
        return torch.zeros_like(z)                     # ar.py:59-66: the reference's reverse is a stub returning zeros

    def logdet(self, input, context=None):
        return self.forward(input, context)[1]
