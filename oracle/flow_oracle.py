"""CPU oracle for the ContextFlow++ flow log-density path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and the CPU-baseline / `--impl reference` legs of bench.py and of its per-row companions
(tools/bench_training.py --impl reference) may import this module.  The product (contextflow_b200/) never does: it fails loudly without its CUDA library.

What it is: a from-scratch restatement of the reference's algorithm for `FlowSequential.log_prob`
as flat functions over a `{state_dict key: tensor}` mapping -- no nn.Module, no einops, no
torch.distributions -- evaluated with torch CPU tensor arithmetic in float32 or float64.  Each function
cites the reference file:line it follows (paths relative to /root/reference/contextflow).

Parity status: PINNED.  The reference publishes no golden vectors for this path, so the
oracle is pinned against outputs of the unmodified reference executed in the build container:
tests/golden/*.npz, produced by tests/golden/make_golden.py (per-layer ldj, per-layer z checksums, final
z and log-prob for every BASELINE configuration and the encoder/context variants), checked by
tests/test_oracle_golden.py.

Third-party arithmetic on the path: the reference is a PyTorch program (torch unpinned in
requirements.txt, conda pins 2.0.1; this image has 2.11.0) and uses einops only for index permutations.
The oracle uses torch CPU kernels for conv2d / matmul / erf / tanh etc., i.e. the same arithmetic substrate.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

# Where the restatement evaluates.  'cpu' always, except bench.py --impl reference --ref-device cuda, which runs the same eager torch op
# sequence on the GPU (inside `with torch.device('cuda')`) to time the reference's torch-on-CUDA path.
DEVICE = 'cpu'
import torch.nn.functional as F

LOG2PI_HALF = 0.5 * math.log(2 * math.pi)
GMM_COMPONENTS = 8          # model.py:115
TS_DATASETS = ('atm', 'msl', 'smd', 'smap')


# =================================================================================================
# model structure: restates create_model (model.py:95-163) and ContextEncoder (model.py:30-90)
# =================================================================================================
def encoder_spec(contexts, enc_emb, enc_type, data_size0) -> dict:
    """ContextEncoder.__init__ (model.py:31-90): embedding kind, width C, surjection kind."""
    n = len(contexts)
    spec = dict(contexts=list(contexts), emb=enc_emb, type=enc_type, num_cats=None, bits=None)
    if enc_emb == 'onehot':
        C = sum(contexts); spec['num_cats'] = [1] * C                       # model.py:33-35
    elif enc_emb == 'eye' and enc_type == 'argmax':
        bits = [int(np.ceil(np.log2(c))) for c in contexts]                  # dequantize.py:190-194
        C = sum(bits); C += C % 2                                            # model.py:37-38
        spec['bits'] = bits; spec['num_cats'] = list(contexts)
    elif enc_emb == 'eye':
        C = n; spec['num_cats'] = list(contexts)                             # model.py:42-44
    elif enc_emb == 'embed':
        C = data_size0 * n; spec['emb_dim'] = data_size0                     # model.py:46-47
    else:
        raise NotImplementedError(f'{enc_emb} is not supported enc-emb!')
    if enc_type in ('vardeq', 'argmax', 'probsample'):
        if C % 2:
            raise NotImplementedError('odd encoder width: the reference inserts an Augment that mis-shapes the inner flow')
        spec['inner_dim'] = 2 * C // n                                       # model.py:75,79,83
    elif enc_type in ('eyesample', 'uniform'):
        pass
    else:
        raise NotImplementedError(f'{enc_type} is not supported enc-type!')
    if enc_type in ('uniform', 'vardeq', 'argmax') and enc_emb == 'embed':
        raise NotImplementedError('embed x uniform/vardeq/argmax: num_cats undefined in the reference (model.py:76-84)')
    spec['C'] = C
    return spec


def build_stack(cfg: dict, data_size, mixtures: int, contexts) -> dict:
    """create_model (model.py:95-163) as a list of layer descriptors; 'key' = FlowSequential index."""
    L: List[dict] = []
    image = cfg['dataset'] in ('mnist', 'cifar10')
    alpha = 1e-4
    if image:                                                                # model.py:97-100
        L += [dict(op='dequant'), dict(op='normalize', t=0.0, s=256.0),
              dict(op='normalize', t=alpha, s=1 / (1 - 2 * alpha)), dict(op='logit')]
    if not (mixtures == 1 or cfg['dist'] == 'gauss'):
        raise NotImplementedError('only the Gaussian-mixture base is on the hot path')
    ts = cfg['dataset'] in TS_DATASETS
    p, krn, pad = ((2, 1), (3, 1), (1, 0)) if ts else ((2, 2), (3, 3), (1, 1))   # model.py:114
    K = GMM_COMPONENTS
    special = not cfg['generalist']
    cf = bool(cfg['contextflow'])
    ee, et = cfg['enc_emb'], cfg['enc_type']

    def enc(d0, emb=ee, typ=et):
        return encoder_spec(contexts, emb, typ, d0) if special else None

    sz = tuple(data_size)
    for l in range(cfg['num_blocks']):
        if sz[0] % 2:                                                        # model.py:121-123
            L.append(dict(op='augment', size=(1, sz[1], sz[2]))); sz = (sz[0] + 1, sz[1], sz[2])
        if cfg['dataset'] not in ('msl', 'smd', 'smap'):                     # model.py:125-127
            L.append(dict(op='squeeze', p=p)); sz = (sz[0] * p[0] * p[1], sz[1] // p[0], sz[2] // p[1])
        for k in range(cfg['block_size']):
            L.append(dict(op='conv1x1', D=sz[0], enc=enc(sz[0]), contextflow=cf))
            if cfg['actnorm']:
                L.append(dict(op='actnorm', D=sz[0], enc=enc(2 * sz[0]), contextflow=cf))
            if cfg['coupling'] == 'trans' and sz[1] % p[0] == 0 and sz[2] % p[1] == 0:
                L.append(dict(op='transcoupling', size=sz, p=p, enc=enc(sz[0]), contextflow=cf))
            elif cfg['coupling'] == 'conv':
                L.append(dict(op='coupling', C=sz[0], krn=krn, pad=pad, enc=enc(sz[0]), contextflow=cf))
            elif cfg['coupling'] == 'maf':                                   # model.py:145-147
                if special and not cf:                                       # ar.py:26,44: the 3D-channel concatenation meets a 2D-channel conv1
                    raise RuntimeError('the reference cannot execute a conventional MaskedCoupling specialist (ar.py:26,44)')
                L.append(dict(op='maf', C=sz[0], krn=krn, pad=pad, enc=enc(sz[0]), contextflow=cf))
            if cfg['dataset'] == 'atm':                                      # model.py:149-151
                L.append(dict(op='permute')); sz = (sz[1], sz[0], sz[2])
        if cfg['split_prior'] and l < cfg['num_blocks'] - 1:                 # model.py:153-158
            sz = (sz[0] // 2, sz[1], sz[2])
            L.append(dict(op='splitprior', size=sz, M=mixtures, K=K,
                          enc=enc(2 * mixtures * K * sz[0] // len(contexts), 'embed', 'eyesample'), contextflow=cf))
    for i, lay in enumerate(L):
        lay['key'] = str(i)
    base = dict(op='gmm', size=sz, M=mixtures, K=K, key='dist',
                enc=enc(2 * mixtures * K * sz[0] // len(contexts), 'embed', 'eyesample'), contextflow=cf)
    return dict(layers=L, base=base, M=mixtures, out_size=sz, data_size=tuple(data_size))


# =================================================================================================
# small helpers
# =================================================================================================
class _P:
    """Parameter view: state_dict keys under a prefix, cast to the oracle dtype."""

    def __init__(self, state: Dict[str, torch.Tensor], dt):
        self.s, self.dt = state, dt

    def __call__(self, key):
        v = self.s[key]
        if v.is_floating_point() and v.requires_grad and torch.is_grad_enabled():   # training direction: autograd over the op sequence
            return v.to(DEVICE, self.dt)
        return v.detach().to(DEVICE, self.dt) if v.is_floating_point() else v.detach().to(DEVICE)

    def has(self, key):
        return key in self.s


def logabsdet(A):                                  # torch.slogdet(NN)[1]: conv1x1.py:43,53
    return torch.linalg.slogdet(A)[1]


def softplus(x):                                   # F.softplus, beta=1, threshold=20
    return F.softplus(x)


# =================================================================================================
# index-only layers (bit exact)
# =================================================================================================
def squeeze(x, p):
    """squeeze.py:10-11  'b c (h p1) (w p2) -> b (c p1 p2) h w'."""
    B, C, H, W = x.shape
    p1, p2 = p
    y = x.reshape(B, C, H // p1, p1, W // p2, p2).permute(0, 1, 3, 5, 2, 4)
    return y.reshape(B, C * p1 * p2, H // p1, W // p2).contiguous()


def permute_chw(x):
    """permute_axes.py:13-14 with permutation (0,2,1,3)."""
    return x.permute(0, 2, 1, 3).contiguous()


def int_to_bits(v, bits):
    """dequantize.py:196-211 integer_to_base(base=2): MSB first."""
    powers = 2 ** torch.arange(bits - 1, -1, -1)
    return (v[..., None] // powers) % 2


# =================================================================================================
# context encoders  (model.py:30-90, _embeddings.py, dequantize.py, flowsequential.py:60-69)
# =================================================================================================
def embed_lookup(P, prefix, ctx):
    """CatEmbeddings.forward (_embeddings.py:265-283), stack=False, no bias."""
    return torch.cat([P(f'{prefix}._embeddings.{i}.weight')[ctx[:, i]] for i in range(ctx.shape[1])], 1)


def mlp3(P, prefix, x, conv=False):
    """Linear/1x1-Conv, ReLU, Linear, ReLU, Linear (coupling.py:26-29 with 1x1 kernels; coupling.py:37)."""
    for j in (0, 2, 4):
        w = P(f'{prefix}.{j}.weight'); b = P(f'{prefix}.{j}.bias')
        w = w.reshape(w.shape[0], -1)
        x = x @ w.t() + b
        if j < 4:
            x = torch.relu(x)
    return x


def coupling_elementwise(x, h):
    """coupling.py:50-66: h -> (t, log_s); z1 = x1*exp(log_s)+t; returns z, sum log_s."""
    Ch = x.shape[1] // 2
    t, r = h[:, :Ch], h[:, Ch:]
    log_s = 2.0 * torch.tanh(r / 2.0)
    z1 = x[:, Ch:] * torch.exp(log_s) + t
    return torch.cat([x[:, :Ch], z1], 1), log_s.flatten(1).sum(-1)


def actnorm_init(x):
    """actnorm.py:28-35: mean and log(unbiased std + 1e-8) over all dims but channel."""
    dims = [i for i in range(x.dim()) if i != 1]
    return torch.mean(x, dim=dims), torch.log(torch.std(x, dim=dims) + 1e-8)


def inner_flow_sample(P, state, prefix, ctx, C, noise, dt):
    """FlowInvSequential.sample (flowsequential.py:60-69) over ConditionalGaussianDistribution.sample
    (gaussian.py:263-270) then 2 x [FC (conv1x1.py:80-96), ActNormFC (actnorm.py:86-102), CouplingFC
    (coupling.py:80-97)].  `prefix` = '<layer>.context_net.1.encoder'."""
    c = embed_lookup(P, f'{prefix}.dist.context_net', ctx)                   # (B, 2C)
    mean, log_scale = c[:, :C], c[:, C:]                                     # 'b (p c) -> p b c'
    eps = noise.randn((ctx.shape[0], C)).to(dt)
    x = mean + log_scale.exp() * eps
    logq = (-LOG2PI_HALF - log_scale - 0.5 * torch.exp(-2 * log_scale) * (x - mean) ** 2).sum(-1)
    for base in (0, 3):
        NN = P(f'{prefix}.{base}.NN')
        x = x @ NN.t(); logq = logq - logabsdet(NN)                          # H=W=1
        kt, kl, ki = f'{prefix}.{base + 1}.NN_t', f'{prefix}.{base + 1}.NN_logs', f'{prefix}.{base + 1}.initialized'
        if int(state[ki]) == 0:                                              # actnorm.py:53 (also in eval)
            m, ls = actnorm_init(x.reshape(-1, C, 1, 1))
            state[kt].copy_(m.to(state[kt].dtype)); state[kl].copy_(ls.to(state[kl].dtype)); state[ki].fill_(1)
        t, logs = P(kt), P(kl)
        x = (x - t) * torch.exp(-logs); logq = logq - logs.sum()
        h = mlp3(P, f'{prefix}.{base + 2}.NN', x[:, :C // 2])
        x, ls = coupling_elementwise(x, h); logq = logq - ls
    return x, logq


def sigmoid_flow(P, prefix, u):
    """activations.py:234-238."""
    T = P(f'{prefix}.sigmoid.temperature')
    x = T * u
    ldj = torch.log(T) - softplus(-x) - softplus(x)
    return torch.sigmoid(x), ldj.sum(-1)


def context_encode(P, state, prefix, spec, ctx, noise, dt):
    """ContextEncoder = Sequential(emb, encoder) (model.py:90) -> (c (B,C), logp_c (B,))."""
    if ctx.dim() != 2:
        raise ValueError('The input must have two dimensions')
    B, n = ctx.shape
    if spec['emb'] == 'onehot':                                              # _embeddings.py:140-150
        x = torch.cat([F.one_hot(ctx[:, i], card) for i, card in enumerate(spec['contexts'])], 1)
    elif spec['emb'] == 'eye':                                               # _embeddings.py:103-109
        x = ctx
    else:
        x = embed_lookup(P, f'{prefix}.0', ctx)
    typ, C = spec['type'], spec['C']
    enc = f'{prefix}.1.encoder'
    if typ == 'eyesample':                                                   # dequantize.py:133-136
        return x.to(dt), torch.zeros(B, dtype=dt)
    if typ == 'uniform':                                                     # dequantize.py:55-63
        u = noise.rand(tuple(x.shape)).to(dt)
        z = (x.to(dt) + u) / P(f'{prefix}.1.qbins')
        ldj = (P(f'{prefix}.1.ldj_per_dim') * x.shape[1]).sum(-1).repeat(B)
        return z, ldj
    u, qu = inner_flow_sample(P, state, enc, ctx, C, noise, dt)
    up, act_ldj = sigmoid_flow(P, f'{prefix}.1', u)
    if typ == 'vardeq':                                                      # dequantize.py:107-116
        z = (x.to(dt) + up) / P(f'{prefix}.1.qbins')
        ldj = (P(f'{prefix}.1.ldj_per_dim') * x.shape[1]).sum(-1).repeat(B)
        return z, ldj + act_ldj - qu
    if typ == 'argmax':                                                      # dequantize.py:239-268
        bits = torch.cat([int_to_bits(ctx[:, i], b) for i, b in enumerate(spec['bits'])], -1)
        if bits.shape[-1] % 2:
            bits = torch.cat([bits, torch.zeros(B, 1, dtype=bits.dtype)], -1)
        sign = (bits * 2 - 1).to(dt)
        return up * sign, act_ldj - qu
    if typ == 'probsample':                                                  # dequantize.py:152-161
        return up, act_ldj + qu
    raise NotImplementedError(typ)


# =================================================================================================
# flow layers
# =================================================================================================
def conv1x1(P, state, lay, x, ctx, noise, dt):
    """Conv1x1.forward (conv1x1.py:28-57)."""
    k, D = lay['key'], lay['D']
    B, _, H, W = x.shape
    NN = P(f'{k}.NN')
    if lay['enc'] is None:
        z = torch.einsum('ij,bjhw->bihw', NN, x)
        return z, (logabsdet(NN) * H * W).expand(B).clone()
    c, logp_c = context_encode(P, state, f'{k}.context_net', lay['enc'], ctx, noise, dt)
    c = (c @ P(f'{k}.CN.weight').t() + P(f'{k}.CN.bias')).reshape(B, D, D)
    c_diag = torch.diagonal(c, dim1=-2, dim2=-1)
    c_ldj = c_diag.sum(-1)
    Wb = torch.tril(c, diagonal=-1) + torch.diag_embed(torch.exp(c_diag))
    if lay['contextflow']:
        Wb = Wb - torch.eye(D, dtype=dt) + NN
        ldj = H * W * (logabsdet(NN) + c_ldj)
    else:
        ldj = H * W * c_ldj
    z = torch.einsum('bij,bjhw->bihw', Wb, x)
    return z, ldj + logp_c * H * W


def actnorm(P, state, lay, x, ctx, noise, dt):
    """ActNorm.forward (actnorm.py:37-60); ldj = +sum(logs) without an H*W factor (App. C-1)."""
    k, D = lay['key'], lay['D']
    B, _, H, W = x.shape

    def base():
        if int(state[f'{k}.initialized']) == 0:
            m, ls = actnorm_init(x)
            state[f'{k}.NN_t'].copy_(m.to(state[f'{k}.NN_t'].dtype))
            state[f'{k}.NN_logs'].copy_(ls.to(state[f'{k}.NN_logs'].dtype))
            state[f'{k}.initialized'].fill_(1)
        return P(f'{k}.NN_t')[None, :].expand(B, D), P(f'{k}.NN_logs')[None, :].expand(B, D)

    if lay['enc'] is None:
        t, logs = base(); logp_c = torch.zeros(B, dtype=dt)
    else:
        c, logp_c = context_encode(P, state, f'{k}.context_net', lay['enc'], ctx, noise, dt)
        c = c @ P(f'{k}.CN.weight').t() + P(f'{k}.CN.bias')                 # (B, 2D): 'b (p d)'
        logp_c = logp_c * H * W
        t, logs = c[:, :D], c[:, D:]
        if lay['contextflow']:
            bt, bl = base(); t = t + bt; logs = logs + bl
    z = (x - t[:, :, None, None]) * torch.exp(-logs[:, :, None, None])
    return z, logs.sum(-1) + logp_c


def conv_conditioner(P, prefix, x0, krn, pad):
    """Coupling.NN (coupling.py:26-29): 1x1 -> ReLU -> kxk reflect -> ReLU -> 1x1."""
    h = torch.relu(F.conv2d(x0, P(f'{prefix}.0.weight'), P(f'{prefix}.0.bias')))
    if pad[0] or pad[1]:
        h = F.pad(h, (pad[1], pad[1], pad[0], pad[0]), mode='reflect')
    h = torch.relu(F.conv2d(h, P(f'{prefix}.2.weight'), P(f'{prefix}.2.bias')))
    return F.conv2d(h, P(f'{prefix}.4.weight'), P(f'{prefix}.4.bias'))


def posemb_sincos_2d(h, w, dim, temperature=10000):
    """simple_vit.py:18-27 (float32 arithmetic, as the reference builds it at construction)."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1).type(torch.float32)


def layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def vit_conditioner(P, prefix, x0, p, T, depth=6, dim_head=64):
    """SimpleViT.forward (simple_vit.py:117-127), heads=1, dim=mlp_dim=T."""
    B, c, Hh, Ww = x0.shape
    p1, p2 = p
    h, w = Hh // p1, Ww // p2
    tok = x0.reshape(B, c, h, p1, w, p2).permute(0, 2, 4, 3, 5, 1).reshape(B, h * w, p1 * p2 * c)
    pe = f'{prefix}.to_patch_embedding'
    x = layer_norm(tok, P(f'{pe}.1.weight'), P(f'{pe}.1.bias'))
    x = x @ P(f'{pe}.2.weight').t() + P(f'{pe}.2.bias')
    x = layer_norm(x, P(f'{pe}.3.weight'), P(f'{pe}.3.bias'))
    x = x + posemb_sincos_2d(h, w, T).to(x.dtype)
    tr = f'{prefix}.transformer'
    for l in range(depth):
        a = f'{tr}.layers.{l}.0'
        y = layer_norm(x, P(f'{a}.norm.weight'), P(f'{a}.norm.bias'))
        qkv = y @ P(f'{a}.to_qkv.weight').t()
        q, k_, v = qkv[..., :dim_head], qkv[..., dim_head:2 * dim_head], qkv[..., 2 * dim_head:]
        dots = (q @ k_.transpose(-1, -2)) * dim_head ** -0.5
        x = (torch.softmax(dots, -1) @ v) @ P(f'{a}.to_out.weight').t() + x
        f = f'{tr}.layers.{l}.1.net'
        y = layer_norm(x, P(f'{f}.0.weight'), P(f'{f}.0.bias'))
        y = F.gelu(y @ P(f'{f}.1.weight').t() + P(f'{f}.1.bias'))
        x = y @ P(f'{f}.3.weight').t() + P(f'{f}.3.bias') + x
    x = layer_norm(x, P(f'{tr}.norm.weight'), P(f'{tr}.norm.bias'))
    cc = T // (p1 * p2)
    return x.reshape(B, h, w, p1, p2, cc).permute(0, 5, 1, 3, 2, 4).reshape(B, cc, h * p1, w * p2)


def coupling(P, state, lay, x, ctx, noise, dt, trans):
    """Coupling.forward / TransCoupling.forward (coupling.py:39-66, 123-148)."""
    k = lay['key']
    B, C, H, W = x.shape
    x0 = x[:, :C // 2]
    if trans:
        T = C * lay['p'][0] * lay['p'][1]
        seq = lay['enc'] is None or lay['contextflow']                       # App. C-6: NN.0.* vs NN.*
        net = lambda inp: vit_conditioner(P, f'{k}.NN.0' if seq else f'{k}.NN', inp, lay['p'], T)
    else:
        net = lambda inp: conv_conditioner(P, f'{k}.NN', inp, lay['krn'], lay['pad'])
    if lay['enc'] is None:
        h = net(x0); logp_c = torch.zeros(B, dtype=dt)
    else:
        c, logp_c = context_encode(P, state, f'{k}.context_net', lay['enc'], ctx, noise, dt)
        if not trans:
            logp_c = logp_c * H * W                                          # coupling.py:43 (not :126)
        cn = mlp3(P, f'{k}.CN', c)
        if lay['contextflow']:
            h = net(x0) + cn[:, :, None, None]
        else:
            h = net(torch.cat([x0, cn[:, :, None, None].expand(B, C, H, W)], 1))
    z, ls = coupling_elementwise(x, h)
    return z, ls + logp_c


def maf_mask(out_c, in_c, kh, kw, data_channels):
    """mask_conv2d(mask_type='B') (layers/autoregressive/utils.py:25-92)."""
    base = torch.ones(data_channels, data_channels).tril(0)
    rows = torch.cat([base] * (in_c // data_channels + 1), 1)
    chan = torch.cat([rows] * (out_c // data_channels + 1), 0)[:out_c, :in_c]
    m = torch.ones(out_c, in_c, kh, kw)
    m[:, :, kh // 2, kw // 2] = chan
    m[:, :, kh // 2, kw // 2 + 1:] = 0
    m[:, :, kh // 2 + 1:] = 0
    return m


def made_degrees(in_degrees, out_features, data_features):
    """MaskedLinear.get_mask_and_degrees, hidden-layer form (masked_linear.py:77-79): mask (out, in), out degrees."""
    max_, min_ = max(1, data_features - 1), min(1, data_features - 1)
    out_degrees = torch.arange(out_features) % max_ + min_
    return (out_degrees[..., None] >= in_degrees).float(), out_degrees


def masked_residual_linear(P, prefix, c, D):
    """MaskedResidualBlockLinear(C, D, D).forward (masked_linear.py:104-128): three pre-activation masked linear layers (none of them
    built with is_output=True, so all use the hidden-layer mask over get_data_degrees = 1..n) plus the identity c, which broadcasts
    only when C == 2 D or C == 1."""
    h = c
    for name in ('linear1', 'linear2', 'linear3'):
        w = P(f'{prefix}.{name}.weight')
        mask, _ = made_degrees(torch.arange(1, w.shape[1] + 1), w.shape[0], D)
        h = F.linear(torch.relu(h), w * mask.to(w), P(f'{prefix}.{name}.bias'))
    return h + c


def masked_coupling(P, lay, x, state=None, ctx=None, noise=None, dt=torch.float32):
    """MaskedCoupling.forward (ar.py:35-57) over MaskedResidualBlock2d (masked_conv_2d.py:81-98): pre-activation convs with
    mask-multiplied weights (:21-23), identity on both halves; z = x * s + t on all channels.  --contextflow specialist (ar.py:39-42):
    h += CN(c) with CN the masked residual linear block, ldj += H W logp_c."""
    k, D = lay['key'], lay['C']
    pad = lay['pad']
    h = x
    for name in ('conv1', 'conv2', 'conv3'):
        w = P(f'{k}.NN.{name}.weight')
        # masked_conv_2d.py:21-23 multiplies weight.DATA by the mask (outside the autograd graph): the forward sees masked weights, the
        # gradient w.r.t. the parameter is the dense one
        w = w + (w * maf_mask(w.shape[0], w.shape[1], w.shape[2], w.shape[3], D).to(w) - w).detach()
        h = torch.relu(h)
        if name == 'conv2' and (pad[0] or pad[1]):
            h = F.pad(h, (pad[1], pad[1], pad[0], pad[0]), mode='reflect')
        h = F.conv2d(h, w, P(f'{k}.NN.{name}.bias'))
    h = h + x.repeat(1, 2, 1, 1)
    logp_c = 0.0
    if lay.get('enc') is not None:
        c, logp_c = context_encode(P, state, f'{k}.context_net', lay['enc'], ctx, noise, dt)
        logp_c = logp_c * x.shape[2] * x.shape[3]                           # ar.py:39
        h = h + masked_residual_linear(P, f'{k}.CN', c, D)[:, :, None, None]
    t, r = h[:, :D], h[:, D:]
    log_s = 2.0 * torch.tanh(r / 2.0)
    return x * torch.exp(log_s) + t, log_s.flatten(1).sum(-1) + logp_c


def gmm_log_prob(P, state, lay, x, ctx, noise, dt):
    """GaussianMixtureDistribution.log_prob (gaussian.py:142-161) via torch.distributions semantics:
    Categorical(probs) -> logits = log(clamp(p/sum p, eps, 1-eps)); MixtureSameFamily: logsumexp_k(
    log_softmax(logits)_k + sum_{dhw} Normal.log_prob)."""
    k, M, K = lay['key'], lay['M'], lay['K']
    B, D, H, W = x.shape
    mean = P(f'{k}.mG')[None]; s = P(f'{k}.sG')[None]
    logp_c = torch.zeros(B, dtype=dt)
    if lay['enc'] is not None:
        c, logp_c = context_encode(P, state, f'{k}.context_net', lay['enc'], ctx, noise, dt)
        c = c.reshape(B, 2, M, K, D, 1, 1)
        logp_c = logp_c * H * W
        mean = mean + c[:, 0]; s = s + c[:, 1]
    scale = softplus(s)
    xx = x[:, None, None]
    comp = (-((xx - mean) ** 2) / (2 * scale ** 2) - scale.log() - math.log(math.sqrt(2 * math.pi))).sum((-3, -2, -1))
    w = torch.softmax(P(f'{k}.wG'), -1)
    w = w / w.sum(-1, keepdim=True)
    eps = torch.finfo(w.dtype).eps
    logits = torch.log(w.clamp(min=eps, max=1 - eps))
    mix = torch.log_softmax(logits, -1)
    return torch.logsumexp(comp + mix[None], -1) + logp_c[:, None]


def logit_ldj(x):
    """transforms.py:11-18."""
    return torch.log(x) - torch.log(1 - x), (-torch.log(x) - torch.log(1 - x)).flatten(1).sum(-1)


# =================================================================================================
# the container: FlowSequential.forward / log_prob (flowsequential.py:18-30)
# =================================================================================================
def forward(stack: dict, state: Dict[str, torch.Tensor], x: torch.Tensor, ctx: Optional[torch.Tensor], noise,
            dtype=torch.float32, trace: Optional[Callable] = None):
    """Returns (z, logp (B,M)).  `noise` provides rand(shape)/randn(shape) in the reference's draw order."""
    dt = dtype
    P = _P(state, dt)
    x = x.detach().to(DEVICE, dt)
    if ctx is not None:
        ctx = ctx.detach().to(DEVICE)
    B, M = x.shape[0], stack['M']
    logdet = torch.zeros(B, M, dtype=dt)
    for lay in stack['layers']:
        op = lay['op']
        if op == 'dequant':                                                  # dequantize.py:14-17, uniform.py:31-34
            x, ldj = x + noise.rand(tuple(x.shape)).to(dt), torch.zeros(B, dtype=dt)
        elif op == 'normalize':                                              # normalize.py:27-49 (scalar scale)
            Cn, Dn = x.shape[1], x[0, 0].numel()
            s = torch.tensor([lay['s']], dtype=torch.float32).to(dt); t = torch.tensor([lay['t']], dtype=torch.float32).to(dt)
            ldj = (Cn * (-1 * Dn * torch.log(s).sum())).expand(B)
            x = x / s + t
        elif op == 'logit':
            x, ldj = logit_ldj(x)
        elif op == 'augment':                                                # augment.py:14-18, gaussian.py:50-72
            e = noise.randn((B,) + tuple(lay['size'])).to(dt)
            logq = (-LOG2PI_HALF - 0.5 * e ** 2).flatten(1).sum(-1).unsqueeze(-1)
            x, ldj = torch.cat([x, e], 1), -logq
        elif op == 'squeeze':
            x, ldj = squeeze(x, lay['p']), torch.zeros(B, dtype=dt)
        elif op == 'permute':
            x, ldj = permute_chw(x), torch.zeros(B, dtype=dt)
        elif op == 'conv1x1':
            x, ldj = conv1x1(P, state, lay, x, ctx, noise, dt)
        elif op == 'actnorm':
            x, ldj = actnorm(P, state, lay, x, ctx, noise, dt)
        elif op == 'coupling':
            x, ldj = coupling(P, state, lay, x, ctx, noise, dt, trans=False)
        elif op == 'transcoupling':
            x, ldj = coupling(P, state, lay, x, ctx, noise, dt, trans=True)
        elif op == 'maf':
            x, ldj = masked_coupling(P, lay, x, state, ctx, noise, dt)
        elif op == 'splitprior':                                             # splitprior.py:12-15
            Ch = x.shape[1] // 2
            ldj = gmm_log_prob(P, state, dict(lay, key=f"{lay['key']}.dist"), x[:, Ch:], ctx, noise, dt)
            x = x[:, :Ch].contiguous()
        else:
            raise NotImplementedError(op)
        logdet = logdet + (ldj if ldj.dim() == 2 else ldj.unsqueeze(-1))     # flowsequential.py:23
        if trace is not None:
            trace(lay, x, ldj)
    logp = gmm_log_prob(P, state, stack['base'], x, ctx, noise, dt)
    return x, logp + logdet


def log_prob(stack, state, x, ctx, noise, dtype=torch.float32):
    return forward(stack, state, x, ctx, noise, dtype)[1]


def bits_per_dim(logp, data_size):
    """-logsumexp_m(logp) / (ln2 * prod(data_size)) (experiment_cl.py:55,127,200 `dim_inv` convention)."""
    return -torch.logsumexp(logp, -1) / (math.log(2) * float(np.prod(data_size)))


# =================================================================================================
# inverse direction: every layer's .reverse as the reference can execute it, and the loop of
# FlowSequential.sample (flowsequential.py:32-39) from a given latent
# =================================================================================================
def unsqueeze(y, p):
    """squeeze.py:13-14  'b (c p1 p2) h w -> b c (h p1) (w p2)'."""
    B, Cs, Hs, Ws = y.shape
    p1, p2 = p
    x = y.reshape(B, Cs // (p1 * p2), p1, p2, Hs, Ws).permute(0, 1, 4, 2, 5, 3)
    return x.reshape(B, Cs // (p1 * p2), Hs * p1, Ws * p2).contiguous()


def coupling_reverse(P, state, lay, z, ctx, noise, dt, trans):
    """Coupling.reverse / TransCoupling.reverse (coupling.py:68-73, 150-155): get_xs_logs_t on z0, x1 = (z1 - t) / s."""
    k = lay['key']
    B, C, H, W = z.shape
    Ch = C // 2
    z0 = z[:, :Ch]
    if trans:
        T = C * lay['p'][0] * lay['p'][1]
        seq = lay['enc'] is None or lay['contextflow']
        net = lambda inp: vit_conditioner(P, f'{k}.NN.0' if seq else f'{k}.NN', inp, lay['p'], T)
    else:
        net = lambda inp: conv_conditioner(P, f'{k}.NN', inp, lay['krn'], lay['pad'])
    if lay['enc'] is None:
        h = net(z0)
    else:
        c, _ = context_encode(P, state, f'{k}.context_net', lay['enc'], ctx, noise, dt)
        cn = mlp3(P, f'{k}.CN', c)
        if lay['contextflow']:
            h = net(z0) + cn[:, :, None, None]
        else:
            h = net(torch.cat([z0, cn[:, :, None, None].expand(B, C, H, W)], 1))
    t, r = h[:, :Ch], h[:, Ch:]
    s = torch.exp(2.0 * torch.tanh(r / 2.0))
    return torch.cat([z0, (z[:, Ch:] - t) / s], 1)


def conv1x1_reverse(P, lay, z):
    """Conv1x1.reverse, context-free branch (conv1x1.py:70): conv with torch.inverse(NN).  The context branch of the reference
    is not executable (reads w_ginv before assignment, :62)."""
    if lay['enc'] is not None:
        raise NotImplementedError('Conv1x1.reverse with context is not executable in the reference')
    return torch.einsum('ij,bjhw->bihw', torch.inverse(P(f"{lay['key']}.NN")), z)


def actnorm_reverse(P, lay, z):
    """ActNorm.reverse, context-free branch (actnorm.py:73-78): x = z * exp(logs) + t."""
    if lay['enc'] is not None:
        raise NotImplementedError('ActNorm.reverse with context is not executable in the reference (actnorm.py:65)')
    k = lay['key']
    return z * torch.exp(P(f'{k}.NN_logs'))[None, :, None, None] + P(f'{k}.NN_t')[None, :, None, None]


def reverse_layer(P, state, lay, z, ctx, noise, dt):
    op = lay['op']
    if op == 'dequant':
        return z.floor()                                                     # dequantize.py:19-20
    if op == 'normalize':                                                    # normalize.py:37-41
        s = torch.tensor([lay['s']], dtype=torch.float32).to(dt); t = torch.tensor([lay['t']], dtype=torch.float32).to(dt)
        return (z - t) * s
    if op == 'logit':
        return torch.sigmoid(z)                                              # transforms.py:14-15
    if op == 'augment':
        return z[:, : z.shape[1] - lay['size'][0]].contiguous()              # augment.py:20-23
    if op == 'squeeze':
        return unsqueeze(z, lay['p'])
    if op == 'permute':
        return permute_chw(z)                                                # (0,2,1,3) is its own inverse
    if op == 'conv1x1':
        return conv1x1_reverse(P, lay, z)
    if op == 'actnorm':
        return actnorm_reverse(P, lay, z)
    if op == 'coupling':
        return coupling_reverse(P, state, lay, z, ctx, noise, dt, trans=False)
    if op == 'transcoupling':
        return coupling_reverse(P, state, lay, z, ctx, noise, dt, trans=True)
    if op == 'splitprior':
        raise AttributeError("'SplitPrior' object has no attribute 'C'")    # splitprior.py:18 in the reference
    raise NotImplementedError(op)


def reverse(stack, state, z, ctx, noise, dtype=torch.float32, trace: Optional[Callable] = None, stop_before: int = 0):
    """The layer loop of FlowSequential.sample (flowsequential.py:34-37) from latent z; layers [stop_before:] are inverted."""
    dt = dtype
    P = _P(state, dt)
    x = z.detach().to(DEVICE, dt)
    if ctx is not None:
        ctx = ctx.detach().to(DEVICE)
    for lay in reversed(stack['layers'][stop_before:]):
        x = reverse_layer(P, state, lay, x, ctx, noise, dt)
        if trace is not None:
            trace(lay, x)
    return x


def gmm_sample_given(P, lay, comp, eps, m=1):
    """The deterministic part of GaussianMixtureDistribution.sample (gaussian.py:163-166): component comp[b] of mixture m = 1,
    x = mG + softplus(sG) * eps."""
    k = lay['key']
    mean = P(f'{k}.mG')[m][comp]; scale = softplus(P(f'{k}.sG')[m][comp])
    return mean + scale * eps


# =================================================================================================
# loss / score epilogue: experiment_ad.py:204-211,262-281, experiment_cl.py:127-133,185-204
# =================================================================================================
def score_epilogue(logp, dim_inv, gt=None, class_w=None):
    s = dim_inv * logp
    s = s.clone(); s[s != s] = 0.0
    lse = torch.logsumexp(s, -1)
    out = dict(scaled=s, lse=lse, softmax1=torch.softmax(s, -1)[:, 1] if s.shape[1] > 1 else torch.ones_like(lse),
               last=s[:, -1], argmax=torch.argmax(s, -1))
    sums = [F.logsigmoid(lse).double().sum(), F.logsigmoid(s).double().sum(), torch.zeros((), dtype=torch.float64),
            torch.zeros((), dtype=torch.float64)]
    if gt is not None:
        w = class_w[gt] if class_w is not None else torch.ones_like(lse)
        sums[2] = (w * (lse - s.gather(1, gt[:, None])[:, 0])).double().sum()
        sums[3] = w.double().sum()
    out['sums'] = torch.stack(sums).float()
    return out


# =================================================================================================
# training direction: the loss of experiment_ad.py:204-209 over `log_prob`; gradients by torch autograd over
# the op sequence above (enable requires_grad on the float tensors of `state`)
# =================================================================================================
def training_loss(logp, gt, data_size, alpha, criterion=True, class_w=None):
    dim_inv = 1.0 / float(np.prod(data_size))
    s = dim_inv * logp
    s = torch.where(s != s, torch.zeros_like(s), s)                          # experiment_ad.py:205
    if criterion:
        uns = -alpha * F.logsigmoid(torch.logsumexp(s, -1)).mean()           # :207
        sup = F.cross_entropy(s, gt, weight=class_w)                         # :208, model.py:294
    else:
        uns = -alpha * F.logsigmoid(s).mean()
        sup = torch.zeros_like(uns)
    return sup + uns, sup, uns


# =================================================================================================
# data format in front of the path: get_windows (datasets/mtad_data_preprocess.py:58-74) + the transpose / float32
# cast of sliding_window_dataset (datasets/mtad_dataloader.py:106-110)
# =================================================================================================
def sliding_windows(ts: np.ndarray, window_size: int, stride: int = 1) -> torch.Tensor:
    rows = ts.shape[0]
    ends = np.arange(0, rows, stride)
    idx = np.maximum(ends[:, None] - window_size + 1 + np.arange(window_size)[None, :], 0)     # replication padding with row 0
    w = np.asarray(ts, dtype=float)[idx]                                                        # (N, L, D) float64
    return torch.tensor(np.transpose(w, (0, 2, 1)), dtype=torch.float).unsqueeze(-1)            # N x D x L x 1


# ---------------------------------------------------------------------------------------------- rows beside the headline path
def sigmoid_layer(x, temperature=1.0):
    """Standalone Sigmoid flow layer, activations.py:234-238: z = sigmoid(T x), ldj = sum_last(log T - softplus(-T x) - softplus(T x))."""
    t = torch.as_tensor(temperature, dtype=x.dtype)
    tx = t * x
    return torch.sigmoid(tx), (torch.log(t) - softplus(-tx) - softplus(tx)).sum(-1)


def sigmoid_layer_reverse(z, temperature=1.0, eps=0.0):
    """activations.py:240-244."""
    zc = torch.clamp(z, eps, 1 - eps)
    return (1.0 / temperature) * (torch.log(zc) - torch.log1p(-zc))


def softplus_layer(x):
    """Standalone Softplus flow layer, activations.py:252-259: z = softplus(x), ldj = sum_last logsigmoid(x)."""
    return softplus(x), (torch.clamp(x, max=0) - torch.log1p(torch.exp(-x.abs()))).sum(-1)


def softplus_layer_reverse(z, eps=1e-7):
    """activations.py:261-264: x = z + log(1 - exp(-max(z, eps)))."""
    return z + torch.log1p(-torch.exp(-z.clamp(eps)))


def student_mixture_log_prob(P, x):
    """StudentMixtureDistribution.log_prob, distributions/student.py:76-98, written out: for both families the mixture weights are
    softmax over dim 0 of w (M, K), renormalised along K by Categorical(probs); component log-densities are summed over (D, H, W);
    log_prob (B, M) = logsumexp_k(Gaussian) + logsumexp_k(Student-t).  StudentT.log_prob as torch.distributions writes it."""
    M, K = P['wG'].shape
    xb = x[:, None, None]                                               # (B, 1, 1, D, H, W) against (M, K, D, H, W)

    def logw(w):
        p = torch.softmax(w, dim=0)
        return torch.log(p / p.sum(-1, keepdim=True))
    sg = softplus(P['sG'])
    lg = (-0.5 * ((xb - P['mG']) / sg) ** 2 - torch.log(sg) - 0.5 * math.log(2 * math.pi)).sum((-1, -2, -3))
    ss, df = softplus(P['sS']), softplus(P['vS'])
    y = (xb - P['mS']) / ss
    Z = torch.log(ss) + 0.5 * torch.log(df) + 0.5 * math.log(math.pi) + torch.lgamma(0.5 * df) - torch.lgamma(0.5 * (df + 1.0))
    ls = (-0.5 * (df + 1.0) * torch.log1p(y ** 2 / df) - Z).sum((-1, -2, -3))
    return torch.logsumexp(lg + logw(P['wG']), -1) + torch.logsumexp(ls + logw(P['wS']), -1)
