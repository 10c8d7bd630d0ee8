"""Training direction of the context-free (generalist) conv stack: torch.autograd.Function wrappers whose forward AND
backward are libcfpp kernels, so that the reference's own training loop (experiment_ad.py:204-213: loss from `model.log_prob`,
`cost_sum.backward()`, the torch optimizer model.py:289 builds) runs unchanged over these layers.

What is covered: Conv1x1, ActNorm, Coupling (conv conditioner), Squeeze, the mixture base and the (B,M) log-det accumulation without a
context_net -- the cfg1 stack.  Context-conditioned (specialist) layers, the ViT conditioner and split priors still raise
NotImplementedError under autograd (next scope row).  Gradients match torch autograd over the reference's op sequence to 1e-4
relative (tests/test_gpu_training.py)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _zeros_like_if_none(g, ref_shape, device):
    return torch.zeros(ref_shape, device=device, dtype=torch.float32) if g is None else g


class SqueezeFn(Function):
    @staticmethod
    def forward(ctx, x, p1, p2):
        ctx.p = (p1, p2)
        return ops.squeeze(x, p1, p2)

    @staticmethod
    def backward(ctx, dy):
        return ops.unsqueeze(dy.contiguous(), *ctx.p), None, None


class UnSqueezeFn(Function):
    @staticmethod
    def forward(ctx, y, p1, p2):
        ctx.p = (p1, p2)
        return ops.unsqueeze(y, p1, p2)

    @staticmethod
    def backward(ctx, dx):
        return ops.squeeze(dx.contiguous(), *ctx.p), None, None


class PermuteFn(Function):
    """PermuteAxes((0,2,1,3)) (permute_axes.py:13-14): the swap is its own inverse."""

    @staticmethod
    def forward(ctx, x):
        return ops.permute_chw(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.permute_chw(dy.contiguous())


class Conv1x1Fn(Function):
    """z = NN x per pixel, ldj = HW log|det NN| (conv1x1.py:52-55)."""

    @staticmethod
    def forward(ctx, x, NN, layer):
        z, ldj = ops.conv1x1(x, NN.detach(), layer.logabsdet())
        ctx.save_for_backward(x, NN)
        ctx.layer = layer
        return z, ldj

    @staticmethod
    def backward(ctx, dz, dldj):
        x, NN = ctx.saved_tensors
        B, D, H, W = x.shape
        dz = _zeros_like_if_none(dz, x.shape, x.device).contiguous()
        w4 = NN.detach().reshape(D, D, 1, 1)
        dNN = None
        if ctx.needs_input_grad[1]:
            dNN, _ = ops.conv2d_bwd_weight(x, D, dz, (D, D, 1, 1), bias=False)
            dNN = dNN.view(D, D)
            if dldj is not None:
                ops.logdet_grad_(dNN, ctx.layer.inverse_matrix(check=False), dldj.contiguous(), H * W)
        dx = ops.conv2d_bwd_data(dz, w4) if ctx.needs_input_grad[0] else None
        return dx, dNN, None


class ActNormFn(Function):
    """z = (x - t) exp(-logs), ldj = sum logs (actnorm.py:50-60, context-free)."""

    @staticmethod
    def forward(ctx, x, t, logs):
        z, ldj = ops.actnorm(x, t.detach(), logs.detach())
        ctx.save_for_backward(x, t, logs)
        return z, ldj

    @staticmethod
    def backward(ctx, dz, dldj):
        x, t, logs = ctx.saved_tensors
        dz = _zeros_like_if_none(dz, x.shape, x.device).contiguous()
        dx, dt, dlogs = ops.actnorm_bwd(x, dz, None if dldj is None else dldj.contiguous(), t.detach(), logs.detach(),
                                        need_dx=ctx.needs_input_grad[0])
        return dx, dt, dlogs


class CouplingConvFn(Function):
    """Coupling with the 3-conv conditioner (coupling.py:26-29,39-66), context-free.  Training forward = three conv launches with the
    post-ReLU activations saved + the coupling kernel; backward = coupling_bwd, then the conditioner's weight / data gradients, the
    last one accumulated onto dx's pass-through half."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3):
        Ch = x.shape[1] // 2
        a1 = ops.conv2d_fwd(x, Ch, w1.detach(), b1.detach(), relu=True)
        a2 = ops.conv2d_fwd(a1, a1.shape[1], w2.detach(), b2.detach(), relu=True)
        h = ops.conv2d_fwd(a2, a2.shape[1], w3.detach(), b3.detach(), relu=False)
        z, ldj = ops.coupling(x, h)
        ctx.save_for_backward(x, a1, a2, h, w1, w2, w3)
        return z, ldj

    @staticmethod
    def backward(ctx, dz, dldj):
        x, a1, a2, h, w1, w2, w3 = ctx.saved_tensors
        Ch = x.shape[1] // 2
        dz = _zeros_like_if_none(dz, x.shape, x.device).contiguous()
        dx, dh = ops.coupling_bwd(x, h, dz, None if dldj is None else dldj.contiguous())
        dw3, db3 = ops.conv2d_bwd_weight(a2, a2.shape[1], dh, w3.shape)
        da2 = ops.conv2d_bwd_data(dh, w3.detach(), act=a2)
        dw2, db2 = ops.conv2d_bwd_weight(a1, a1.shape[1], da2, w2.shape)
        da1 = ops.conv2d_bwd_data(da2, w2.detach(), act=a1)
        dw1, db1 = ops.conv2d_bwd_weight(x, Ch, da1, w1.shape)
        ops.conv2d_bwd_data(da1, w1.detach(), out=dx, accumulate=True)       # dx[:, :Ch] += W1^T da1
        return (dx if ctx.needs_input_grad[0] else None), dw1, db1, dw2, db2, dw3, db3


class GmmFn(Function):
    """GaussianMixtureDistribution.log_prob (gaussian.py:142-161), context-free."""

    @staticmethod
    def forward(ctx, x, mG, sG, wG, layer):
        inv_var, cst = layer._tables.get('train', [sG, wG], lambda: ops.gmm_train_prep(sG.detach(), wG.detach()))
        logp, resp = ops.gmm_train_fwd(x, mG.detach(), inv_var, cst)
        ctx.save_for_backward(x, mG, sG, wG, inv_var, resp)
        return logp

    @staticmethod
    def backward(ctx, g):
        x, mG, sG, wG, inv_var, resp = ctx.saved_tensors
        dx, dmG, dsG, dwG = ops.gmm_train_bwd(x, mG.detach(), sG.detach(), wG.detach(), inv_var, resp, g.contiguous(),
                                              need_dx=ctx.needs_input_grad[0])
        return dx, dmG, dsG, dwG, None


class SplitPriorFn(Function):
    """SplitPrior.forward (splitprior.py:12-15) with a context-free mixture: x -> (x[:, :C/2], log_prob(x[:, C/2:]))."""

    @staticmethod
    def forward(ctx, x, mG, sG, wG, layer):
        half = x.shape[1] // 2
        inv_var, cst = layer._tables.get('train', [sG, wG], lambda: ops.gmm_train_prep(sG.detach(), wG.detach()))
        logp, resp = ops.gmm_train_fwd(x[:, half:], mG.detach(), inv_var, cst)
        ctx.save_for_backward(x, mG, sG, wG, inv_var, resp)
        return ops.slice_channels(x, 0, half), logp

    @staticmethod
    def backward(ctx, dx0, g):
        x, mG, sG, wG, inv_var, resp = ctx.saved_tensors
        half = x.shape[1] // 2
        g = torch.zeros((x.shape[0], mG.shape[0]), device=x.device) if g is None else g.contiguous()
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty_like(x) if need_dx else None
        if need_dx:
            ops.place_channels(_zeros_like_if_none(dx0, (x.shape[0], half) + tuple(x.shape[2:]), x.device).contiguous(), dx, 0)
        _, dmG, dsG, dwG = ops.gmm_train_bwd(x[:, half:], mG.detach(), sG.detach(), wG.detach(), inv_var, resp, g, need_dx=need_dx,
                                             dx_out=dx[:, half:] if need_dx else None)
        return dx, dmG, dsG, dwG, None


class LdjSumFn(Function):
    """logprob + ((0 + ldj_0) + ldj_1) + ...  (flowsequential.py:20-27); a (B,) / (B,1) term receives the row sum of the gradient."""

    @staticmethod
    def forward(ctx, last, M, *terms):
        B = last.shape[0]
        ctx.shapes = [tuple(t.shape) for t in terms]
        ctx.M = M
        return ops.ldj_sum([t.detach() for t in terms], B, M, last.device, last=last.detach())

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        rs = None
        grads = []
        for i, shp in enumerate(ctx.shapes):
            if not ctx.needs_input_grad[2 + i]:
                grads.append(None)
            elif len(shp) == 2 and shp[1] == ctx.M and ctx.M > 1:
                grads.append(g)
            else:
                if rs is None:
                    rs = ops.rowsum(g)
                grads.append(rs.view(shp))
        return (g, None, *grads)


class CouplingVitFn(Function):
    """TransCoupling with the SimpleViT conditioner (coupling.py:100-148, simple_vit.py:30-127), context-free.  The training forward is
    the reference's module stack op by op on token rows (LayerNorm, Linear, single-head attention, GELU, residuals) with the
    activations saved; the backward walks it in reverse.  Parameter order = SimpleViT._sources()."""

    @staticmethod
    def forward(ctx, x, vit, *params):
        g = vit.geom
        c, p1, p2, T, n, depth = g['Cin'], g['p1'], g['p2'], g['T'], g['n_tok'], g['depth']
        B, C, H, W = x.shape
        P = [p.detach() for p in params]
        ln0w, ln0b, pew, peb, ln1w, ln1b, lnfw, lnfb = P[:8]
        tok = ops.patchify(x, c, p1, p2)
        y0, m0, r0 = ops.layernorm_fwd(tok, ln0w, ln0b)
        e = ops.rows_linear(y0, pew, peb)
        X, m1, r1 = ops.layernorm_fwd(e, ln1w, ln1b)
        ops.add_pos_(X, vit.pos_embedding.to(x.device, torch.float32).contiguous(), n)
        saved = [tok, m0, r0, y0, e, m1, r1]
        for l in range(depth):
            anw, anb, qkvw, outw, f0w, f0b, f1w, f1b, f3w, f3b = P[8 + 10 * l: 18 + 10 * l]
            y, ma, ra = ops.layernorm_fwd(X, anw, anb)
            qkv = ops.rows_linear(y, qkvw)
            O, Pm = ops.attention_fwd(qkv, B, n)
            X1 = ops.add(ops.rows_linear(O, outw), X)
            y2, mf, rf = ops.layernorm_fwd(X1, f0w, f0b)
            hpre = ops.rows_linear(y2, f1w, f1b)
            gact = ops.gelu_fwd(hpre)
            X2 = ops.add(ops.rows_linear(gact, f3w, f3b), X1)
            saved += [X, ma, ra, y, qkv, Pm, O, X1, mf, rf, y2, hpre, gact]
            X = X2
        Xf, mz, rz = ops.layernorm_fwd(X, lnfw, lnfb)
        cc = T // (p1 * p2)
        h = ops.patchify_inv(Xf, cc, H, W, p1, p2)
        z, ldj = ops.coupling(x, h)
        saved += [X, mz, rz, x, h]
        ctx.save_for_backward(*saved, *params)
        ctx.meta = (c, p1, p2, T, n, depth, B, H, W, cc, len(saved))
        return z, ldj

    @staticmethod
    def backward(ctx, dz, dldj):
        c, p1, p2, T, n, depth, B, H, W, cc, ns = ctx.meta
        S = ctx.saved_tensors[:ns]
        P = [p.detach() for p in ctx.saved_tensors[ns:]]
        ln0w, ln0b, pew, peb, ln1w, ln1b, lnfw, lnfb = P[:8]
        Xlast, mz, rz, x, h = S[-5:]
        dz = _zeros_like_if_none(dz, x.shape, x.device).contiguous()
        dx, dh = ops.coupling_bwd(x, h, dz, None if dldj is None else dldj.contiguous())
        grads = [None] * len(P)
        dXf = ops.patchify(dh, cc, p1, p2)
        dX, grads[6], grads[7] = ops.layernorm_bwd(Xlast, dXf, lnfw, mz, rz)
        for l in reversed(range(depth)):
            X, ma, ra, y, qkv, Pm, O, X1, mf, rf, y2, hpre, gact = S[7 + 13 * l: 20 + 13 * l]
            o = 8 + 10 * l
            anw, anb, qkvw, outw, f0w, f0b, f1w, f1b, f3w, f3b = P[o: o + 10]
            grads[o + 8], grads[o + 9] = ops.rows_linear_bwd_weight(gact, dX)
            dhpre = ops.gelu_bwd(hpre, ops.rows_linear_bwd_data(dX, f3w))
            grads[o + 6], grads[o + 7] = ops.rows_linear_bwd_weight(y2, dhpre)
            d1, grads[o + 4], grads[o + 5] = ops.layernorm_bwd(X1, ops.rows_linear_bwd_data(dhpre, f1w), f0w, mf, rf)
            dX1 = ops.add(dX, d1)
            grads[o + 3], _ = ops.rows_linear_bwd_weight(O, dX1, bias=False)
            dqkv = ops.attention_bwd(qkv, Pm, ops.rows_linear_bwd_data(dX1, outw), B, n)
            grads[o + 2], _ = ops.rows_linear_bwd_weight(y, dqkv, bias=False)
            d0, grads[o], grads[o + 1] = ops.layernorm_bwd(X, ops.rows_linear_bwd_data(dqkv, qkvw), anw, ma, ra)
            dX = ops.add(dX1, d0)
        tok, m0, r0, y0, e, m1, r1 = S[:7]
        de, grads[4], grads[5] = ops.layernorm_bwd(e, dX, ln1w, m1, r1)
        grads[2], grads[3] = ops.rows_linear_bwd_weight(y0, de)
        dtok, grads[0], grads[1] = ops.layernorm_bwd(tok, ops.rows_linear_bwd_data(de, pew), ln0w, m0, r0)
        ops.patchify_inv(dtok, c, H, W, p1, p2, out=dx, accumulate=True)          # dx[:, :c] += the conditioner's input gradient
        return (dx if ctx.needs_input_grad[0] else None, None, *grads)
