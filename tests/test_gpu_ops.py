"""Op-level GPU tests through the C ABI: edge shapes (ragged / non-multiple-of-4 pixel counts, large D, B=1), both GMM context
paths, index-op round trips, standalone module calls.  Reference values come from the CPU oracle's functions."""
import math
import pytest
import torch

from contextflow_b200 import layers as L, ops, rng, synth
from contextflow_b200.layers.rtdl.nn._embeddings import CatEmbeddings, OneHotEncoder
from oracle import flow_oracle as O
from tests.helpers import L_ATOL, L_RTOL, Z_ATOL, Z_RTOL, assert_close

pytestmark = pytest.mark.gpu
dev = 'cuda'


@pytest.mark.parametrize('B,C,H,W,p', [(3, 2, 4, 6, (2, 2)), (2, 38, 144, 1, (2, 1)), (1, 3, 32, 32, (2, 2)), (5, 4, 14, 14, (2, 2))])
def test_squeeze_bit_exact_and_roundtrip(B, C, H, W, p):
    x = synth.uniform('sq', (B, C, H, W))
    y = ops.squeeze(x.to(dev), *p)
    assert torch.equal(y.cpu(), O.squeeze(x, p))
    assert torch.equal(ops.unsqueeze(y, *p).cpu(), x)


@pytest.mark.parametrize('B,C,H,W', [(2, 76, 72, 1), (3, 5, 7, 3), (1, 33, 65, 1)])
def test_permute_bit_exact(B, C, H, W):
    x = synth.uniform('pm', (B, C, H, W))
    assert torch.equal(ops.permute_chw(x.to(dev)).cpu(), x.permute(0, 2, 1, 3).contiguous())


@pytest.mark.parametrize('B,D,HW,ctx', [(5, 8, 256, False), (7, 64, 16, True), (3, 76, 72, True), (2, 100, 49, False), (9, 26, 8, False),
                                        (4, 12, 49, True), (1, 16, 256, True)])
@pytest.mark.parametrize('contextflow', [True, False])
def test_conv1x1_shapes(B, D, HW, ctx, contextflow):
    x = synth.uniform('c1x', (B, D, HW, 1))
    NN = torch.eye(D) + synth.uniform('c1w', (D, D)) * 0.2 / math.sqrt(D)
    c = synth.uniform('c1c', (B, D * D)) * 0.1 if ctx else None
    lp = synth.uniform('c1l', (B,)) if ctx else None
    lad = ops.slogdet(NN.to(dev))
    assert_close(lad.cpu().numpy(), torch.linalg.slogdet(NN.double())[1].float().reshape(1).numpy(), 1e-5, 1e-6, 'slogdet')
    z, ldj = ops.conv1x1(x.to(dev), NN.to(dev), lad, None if c is None else c.to(dev), None if lp is None else lp.to(dev), contextflow)
    if ctx:
        cm = c.reshape(B, D, D)
        diag = torch.diagonal(cm, dim1=-2, dim2=-1)
        Wb = torch.tril(cm, -1) + torch.diag_embed(torch.exp(diag))
        want_l = HW * diag.sum(-1)
        if contextflow:
            Wb = Wb - torch.eye(D) + NN
            want_l = HW * (torch.linalg.slogdet(NN)[1] + diag.sum(-1))
        want_l = want_l + lp * HW
        want = torch.einsum('bij,bjhw->bihw', Wb, x)
    else:
        want = torch.einsum('ij,bjhw->bihw', NN, x)
        want_l = (torch.linalg.slogdet(NN)[1] * HW).expand(B)
    assert_close(z.cpu().numpy(), want.numpy(), Z_RTOL, Z_ATOL, 'conv1x1 z')
    assert_close(ldj.cpu().numpy(), want_l.numpy(), L_RTOL, L_ATOL, 'conv1x1 ldj')


@pytest.mark.parametrize('B,D,HW,K', [(37, 16, 256, 20), (300, 64, 16, 20), (1000, 32, 64, 20), (9, 76, 72, 8), (5, 72, 76, 8), (13, 36, 76, 8), (21, 12, 16, 3),
                                      (1, 64, 16, 20), (70, 128, 4, 64)])
@pytest.mark.parametrize('contextflow', [True, False])
@pytest.mark.parametrize('an_mode', [2, 1, 0])
def test_conv1x1_with_fused_context_network(B, D, HW, K, contextflow, an_mode):
    """cfpp_conv1x1_ctx_fwd (the CN linear layer + W_b assembly + 1x1 conv + ActNorm in one kernel) against the fp64 restatement of
    conv1x1.py:31-50 / actnorm.py:37-60, on ragged batches that leave partial sample groups."""
    assert ops.conv1x1_ctx_supported(B, D, HW, K)
    x = synth.uniform('f1x', (B, D, HW, 1)) * 2
    NN = torch.eye(D) + synth.uniform('f1w', (D, D)) * 0.2 / math.sqrt(D)
    e = synth.uniform('f1e', (B, K)) * 2
    cw = synth.uniform('f1cw', (D * D, K)) * 0.1
    cb = synth.uniform('f1cb', (D * D,)) * 0.1
    lp = synth.uniform('f1l', (B,))
    tl = synth.uniform('f1t', (B, 2 * D)) * 0.5
    alp = synth.uniform('f1al', (B,))
    lad = ops.slogdet(NN.to(dev))
    wt, bt = ops.pack_cn_tril(cw.to(dev), cb.to(dev), D)
    kw = {}
    if an_mode == 2:
        kw = dict(an_t=tl.to(dev), an_logs=None, an_logp_c=alp.to(dev), an_logp_scale=float(HW))
    elif an_mode == 1:
        kw = dict(an_t=tl[:, :D].contiguous().to(dev), an_logs=tl[:, D:].contiguous().to(dev))
    z, ldj = ops.conv1x1_ctx(x.to(dev), e.to(dev), wt, bt, NN.to(dev), lad, lp.to(dev), contextflow, **kw)
    xd, ed = x.double(), e.double()
    cm = (ed @ cw.double().t() + cb.double()).reshape(B, D, D)
    diag = torch.diagonal(cm, dim1=-2, dim2=-1)
    Wb = torch.tril(cm, -1) + torch.diag_embed(torch.exp(diag))
    want_l = HW * diag.sum(-1)
    if contextflow:
        Wb = Wb - torch.eye(D, dtype=torch.float64) + NN.double()
        want_l = HW * (torch.linalg.slogdet(NN.double())[1] + diag.sum(-1))
    want_l = want_l + lp.double() * HW
    want = torch.einsum('bij,bjhw->bihw', Wb, xd)
    if an_mode:
        t, logs = tl[:, :D].double(), tl[:, D:].double()
        want = (want - t[:, :, None, None]) * torch.exp(-logs)[:, :, None, None]
        want_l = want_l + logs.sum(-1) + (HW * alp.double() if an_mode == 2 else 0)
    assert_close(z.cpu().numpy(), want.float().numpy(), Z_RTOL, Z_ATOL * 4, 'conv1x1_ctx z')
    assert_close(ldj.cpu().numpy(), want_l.float().numpy(), L_RTOL, L_ATOL, 'conv1x1_ctx ldj')
    # and the two-kernel route (cn_batch-style linear + cfpp_conv1x1_fwd) gives the same layer
    c2 = ops.linear(e.to(dev), ops.pack_kmajor(cw.to(dev), 1), cb.to(dev))
    z2, l2 = ops.conv1x1(x.to(dev), NN.to(dev), lad, c2, lp.to(dev), contextflow, **kw)
    assert_close(z.cpu().numpy(), z2.cpu().numpy(), Z_RTOL, Z_ATOL * 4, 'fused vs two-kernel z')
    assert_close(ldj.cpu().numpy(), l2.cpu().numpy(), L_RTOL, L_ATOL, 'fused vs two-kernel ldj')


@pytest.mark.parametrize('B,C,HW', [(3, 16, 256), (5, 26, 8), (2, 8, 49), (1, 76, 72), (70, 2, 1)])
@pytest.mark.parametrize('with_ctx', [False, True])
def test_coupling_elementwise(B, C, HW, with_ctx):
    x = synth.uniform('cpx', (B, C, HW, 1)) * 2
    h = synth.uniform('cph', (B, C, HW, 1)) * 3
    add = synth.uniform('cpa', (B, C)) if with_ctx else None
    lp = synth.uniform('cpl', (B,)) if with_ctx else None
    z, ldj = ops.coupling(x.to(dev), h.to(dev), None if add is None else add.to(dev), None if lp is None else lp.to(dev), 3.0)
    hh = h + add[:, :, None, None] if with_ctx else h
    zo, lo = O.coupling_elementwise(x, hh)
    if with_ctx:
        lo = lo + 3.0 * lp
    assert torch.equal(z.cpu()[:, :C // 2], x[:, :C // 2]), 'pass-through half must be copied bit exactly'
    assert_close(z.cpu().numpy(), zo.numpy(), Z_RTOL, Z_ATOL, 'coupling z')
    assert_close(ldj.cpu().numpy(), lo.numpy(), L_RTOL, L_ATOL, 'coupling ldj')


def _gmm_oracle(x, mG, sG, wG, c, M, K):
    state = {'g.mG': mG, 'g.sG': sG, 'g.wG': wG}
    lay = dict(key='g', M=M, K=K, enc=None)
    if c is None:
        return O.gmm_log_prob(O._P(state, torch.float64), state, lay, x.double(), None, None, torch.float64)
    B, D = x.shape[0], x.shape[1]
    cc = c.double().reshape(B, 2, M, K, D, 1, 1)
    mean = mG.double()[None] + cc[:, 0]; scale = torch.nn.functional.softplus(sG.double()[None] + cc[:, 1])
    xx = x.double()[:, None, None]
    comp = (-((xx - mean) ** 2) / (2 * scale ** 2) - scale.log() - math.log(math.sqrt(2 * math.pi))).sum((-3, -2, -1))
    mix = torch.log_softmax(wG.double(), -1)
    return torch.logsumexp(comp + mix[None], -1)


@pytest.mark.parametrize('B,M,K,D,H,W,cards', [(11, 10, 8, 8, 16, 16, [15, 5]), (6, 2, 8, 38, 36, 1, [68]), (19, 1, 8, 26, 8, 1, [7]),
                                                 (3, 3, 8, 4, 7, 7, [4, 3]), (1, 10, 8, 64, 4, 4, [15, 5]), (300, 10, 8, 16, 8, 8, [15, 5]), (150, 5, 3, 6, 5, 3, [9])])
def test_gmm_all_paths(B, M, K, D, H, W, cards):
    x = synth.uniform('gx', (B, D, H, W)) * 2
    mG, sG, wG = synth.uniform('gm', (M, K, D, H, W)), 1 + 0.3 * synth.uniform('gs', (M, K, D, H, W)), synth.uniform('gw', (M, K))
    n = len(cards)
    width = 2 * M * K * D // n
    tabs = [synth.uniform(f'gt{i}', (card, width)) * 0.2 for i, card in enumerate(cards)]
    ctx = torch.stack([synth.randint(f'gc{i}', (B,), card) for i, card in enumerate(cards)], 1)
    c = torch.cat([t[ctx[:, i]] for i, t in enumerate(tabs)], 1)
    want0 = _gmm_oracle(x, mG, sG, wG, None, M, K)
    want1 = _gmm_oracle(x, mG, sG, wG, c, M, K)
    cu = lambda t: t.to(dev)
    got0 = ops.gmm_logprob(cu(x), cu(mG), cu(sG), cu(wG))
    assert_close(got0.cpu().numpy(), want0.numpy(), L_RTOL, L_ATOL, 'gmm no-context')
    c_dev = ops.embed_lookup(cu(ctx), [cu(t) for t in tabs])
    assert torch.equal(c_dev.cpu(), c), 'embedding lookup is an exact gather'
    got1 = ops.gmm_logprob(cu(x), cu(mG), cu(sG), cu(wG), c_dev)
    assert_close(got1.cpu().numpy(), want1.numpy(), L_RTOL, L_ATOL, 'gmm per-sample context')
    got2 = ops.gmm_logprob_ctxtab(cu(x), cu(mG), cu(sG), cu(wG), cu(ctx), cards, [cu(t) for t in tabs])
    assert got2 is not None
    assert_close(got2.cpu().numpy(), want1.numpy(), L_RTOL, L_ATOL, 'gmm bucketed context tables')
    # in-place read of a channel-slice view (SplitPrior): same numbers from x embedded as the second half of a wider tensor
    wide = torch.cat([torch.zeros_like(x), x], 1).to(dev)
    got3 = ops.gmm_logprob(wide[:, D:], cu(mG), cu(sG), cu(wG))
    assert torch.equal(got3, got0)
    # register-tiled kernel (csrc/gmm_tile.cu): no context, bucketed scale context + per-sample mean offsets, strided view
    tab0 = ops.gmm_tile_table(cu(mG), cu(sG), cu(wG))
    assert tab0 is not None
    assert_close(ops.gmm_tile_logprob(cu(x), tab0, M, K).cpu().numpy(), want0.numpy(), L_RTOL, L_ATOL, 'gmm tile no-context')
    assert_close(ops.gmm_tile_logprob(wide[:, D:], tab0, M, K).cpu().numpy(), want0.numpy(), L_RTOL, L_ATOL, 'gmm tile strided view')
    sf, soff = n - 1, (0 if n == 2 else M * K * D)
    tab1 = ops.gmm_tile_table(cu(mG), cu(sG), cu(wG), cu(tabs[sf]), soff)
    got4 = ops.gmm_tile_logprob(cu(x), tab1, M, K, cu(ctx), cards, cu(tabs[0]), 0)
    assert_close(got4.cpu().numpy(), want1.numpy(), L_RTOL, L_ATOL, 'gmm tile context tables')
    if (D * H * W) % 4 == 0:                                      # misaligned view: element-wise cp.async path
        flat = torch.zeros(B * D * H * W + 1, device=dev); flat[1:] = cu(x).flatten()
        assert_close(ops.gmm_tile_logprob(flat[1:].view(B, D, H, W), tab0, M, K).cpu().numpy(), want0.numpy(), L_RTOL, L_ATOL, 'gmm tile misaligned x')


def test_actnorm_stats_and_modes():
    B, D, HW = 6, 10, 12
    x = synth.uniform('anx', (B, D, HW, 1)) * 3 + 1
    m, ls = ops.actnorm_stats(x.to(dev))
    mo, lo = O.actnorm_init(x)
    assert_close(m.cpu().numpy(), mo.numpy(), 1e-5, 1e-6, 'mean'); assert_close(ls.cpu().numpy(), lo.numpy(), 1e-5, 1e-6, 'logstd')
    c = synth.uniform('anc', (B, 2 * D)); lp = synth.uniform('anl', (B,))
    for mode in (0, 1, 2):
        z, ldj = ops.actnorm(x.to(dev), m, ls, c.to(dev) if mode else None, lp.to(dev) if mode else None, 5.0, mode)
        t = (mo if mode != 2 else 0) + (c[:, :D] if mode else 0)
        lg = (lo if mode != 2 else 0) + (c[:, D:] if mode else 0)
        t = t.expand(B, D) if torch.is_tensor(t) else t; lg = lg.expand(B, D)
        want = (x - t[:, :, None, None]) * torch.exp(-lg[:, :, None, None])
        assert_close(z.cpu().numpy(), want.numpy(), Z_RTOL, Z_ATOL, f'actnorm z mode {mode}')
        assert_close(ldj.cpu().numpy(), (lg.sum(-1) + (5.0 * lp if mode else 0)).numpy(), L_RTOL, L_ATOL, f'actnorm ldj mode {mode}')


def test_standalone_embedding_and_surjection_modules():
    ctx = torch.stack([synth.randint('sc0', (9,), 6), synth.randint('sc1', (9,), 3)], 1)
    oh, c2 = OneHotEncoder([6, 3]).to(dev)(ctx.to(dev))
    want = torch.cat([torch.nn.functional.one_hot(ctx[:, 0], 6), torch.nn.functional.one_hot(ctx[:, 1], 3)], 1)
    assert oh.dtype == torch.int64 and torch.equal(oh.cpu(), want) and torch.equal(c2.cpu(), ctx)
    emb = CatEmbeddings([6, 3], 5).to(dev)
    e, _ = emb(ctx.to(dev))
    assert torch.equal(e.cpu(), torch.cat([emb._embeddings[i].weight.detach().cpu()[ctx[:, i]] for i in range(2)], 1))
    surj = L.UniformCatDequantization(num_cats=[6, 3]).to(dev)
    tape = synth.NoiseTape('su')
    with rng.use_source(tape):
        z, ldj = surj((ctx.to(dev), ctx.to(dev)))
    u = synth.NoiseTape('su').rand((9, 2))
    assert_close(z.cpu().numpy(), ((ctx.float() + u) / torch.tensor([6.0, 3.0])).numpy(), 1e-6, 1e-7, 'uniform dequant')
    assert_close(ldj.cpu().numpy(), torch.full((9,), float(-2 * (math.log(6) + math.log(3)))).numpy(), 1e-6, 1e-6, 'uniform dequant ldj (x num_dims quirk)')
    with pytest.raises(ValueError):
        OneHotEncoder([6, 3]).to(dev)(ctx[:, 0].to(dev))


def test_linear_and_ldj_accumulate():
    x = synth.uniform('lx', (70, 20)); w = synth.uniform('lw', (33, 20)); b = synth.uniform('lb', (33,))
    y = ops.linear(x.to(dev), ops.pack_kmajor(w.to(dev), 1), b.to(dev), relu=True)
    assert_close(y.cpu().numpy(), torch.relu(x @ w.t() + b).numpy(), 1e-5, 1e-5, 'linear')
    ld = torch.zeros(5, 3, device=dev)
    ops.ldj_accumulate(ld, torch.arange(5., device=dev)); ops.ldj_accumulate(ld, torch.ones(5, 1, device=dev)); ops.ldj_accumulate(ld, torch.ones(5, 3, device=dev))
    assert torch.equal(ld.cpu(), (torch.arange(5.)[:, None] + 2).expand(5, 3))


@pytest.mark.parametrize('B', [1, 16, 45])
def test_cn_batch_matches_linear_chain(B):
    """cfpp_cn_batch: heterogeneous chains (3-layer ReLU MLP, single Linear, Linear into a lower-triangular D x D matrix) in one
    launch, against fp64 torch on the CPU."""
    specs = [dict(K=20, N=[32, 32, 16], tril=0), dict(K=20, N=[32], tril=0), dict(K=20, N=[64 * 64], tril=64), dict(K=8, N=[152, 152, 76], tril=0),
             dict(K=7, N=[9], tril=3), dict(K=20, N=[256, 256, 128], tril=0), dict(K=5, N=[5 * 5], tril=5)]
    jobs, ins, keep, want = [], [], [], []
    for i, s in enumerate(specs):
        x = synth.uniform(f'cnx{i}', (B, s['K'])) * 2 - 1
        layers, h, K = [], x.double(), s['K']
        for l, N in enumerate(s['N']):
            w = (synth.uniform(f'cnw{i}{l}', (N, K)) * 2 - 1) / math.sqrt(K)
            b = synth.uniform(f'cnb{i}{l}', (N,)) - 0.5 if (i + l) % 3 else None
            layers.append((ops.pack_kmajor(w.to(dev), 1), None if b is None else b.to(dev)))
            h = h @ w.double().t() + (0 if b is None else b.double())
            if l + 1 < len(s['N']):
                h = torch.relu(h)
            K = N
        j, k = ops.cn_job(layers, s['tril'])
        jobs.append(j); keep.append(k); ins.append(x.to(dev)); want.append(h)
    outs = ops.cn_batch(jobs, ins)
    torch.cuda.synchronize()
    for s, o, w in zip(specs, outs, want):
        o = o.cpu().double()
        if s['tril']:
            D = s['tril']
            mask = torch.tril(torch.ones(D, D, dtype=torch.bool)).reshape(-1)
            o, w = o[:, mask], w[:, mask]
        assert_close(o.numpy(), w.numpy(), 1e-5, 1e-5, f'cn_batch {s}')


def test_ldj_sum_is_the_ordered_chain():
    B, M = 13, 10
    terms = [synth.uniform('ls0', (B,)) * 100, synth.uniform('ls1', (B, 1)) * 1e-3, synth.uniform('ls2', (B, M)), synth.uniform('ls3', (B,)) * 7]
    last = synth.uniform('ls4', (B, M)) * 1000
    want = torch.zeros(B, M)
    for t in terms:
        want += t if t.dim() == 2 else t.unsqueeze(-1)
    want = last + want
    got = ops.ldj_sum([t.to(dev) for t in terms], B, M, dev, last=last.to(dev))
    assert torch.equal(got.cpu(), want)                          # same order of fp32 additions: bit identical
    many = [synth.uniform(f'lm{i}', (B,)) for i in range(150)]   # more than one launch worth of terms
    want = torch.zeros(B, M)
    for t in many:
        want += t.unsqueeze(-1)
    assert torch.equal(ops.ldj_sum([t.to(dev) for t in many], B, M, dev).cpu(), want)
    assert torch.equal(ops.ldj_sum([], B, M, dev).cpu(), torch.zeros(B, M))


# ---- rows beside the headline path: standalone activations, Student-t mixture -------------------------------------------------------
def _extras():
    import os
    import numpy as np
    from tests.helpers import GOLD
    return dict(np.load(os.path.join(GOLD, 'extras.npz'), allow_pickle=False))


def test_activation_layers_match_reference_fixture():
    g = _extras()
    x = torch.from_numpy(g['act_x']).to(dev)
    for T in (1.0, 2.5):
        lay = L.Sigmoid(temperature=T, eps=1e-6).to(dev)
        z, ldj = lay(x)
        assert_close(z.cpu().numpy(), g[f'sig_z_{T}'], 1e-6, 1e-6, f'sigmoid z T={T}')
        assert_close(ldj.cpu().numpy(), g[f'sig_ldj_{T}'], 1e-5, 1e-4, f'sigmoid ldj T={T}')
        # the reverse is ill-conditioned where z rounds to 1 (x = 25): compare where the reference's own z is away from the clamp
        rev, ref = lay.reverse(z).cpu(), torch.from_numpy(g[f'sig_rev_{T}'])
        ok = (torch.from_numpy(g[f'sig_z_{T}']) < 1 - 1e-4) & (torch.from_numpy(g[f'sig_z_{T}']) > 1e-4)
        assert_close(rev[ok].numpy(), ref[ok].numpy(), 2e-3, 2e-3, f'sigmoid reverse T={T}')
    lay = L.Softplus()
    z, ldj = lay(x)
    assert_close(z.cpu().numpy(), g['sp_z'], 1e-6, 1e-6, 'softplus z')
    assert_close(ldj.cpu().numpy(), g['sp_ldj'], 1e-5, 1e-4, 'softplus ldj')
    zr = torch.from_numpy(g['sp_z'])
    ok = zr > 1e-3
    assert_close(lay.reverse(z).cpu()[ok].numpy(), torch.from_numpy(g['sp_rev'])[ok].numpy(), 1e-4, 1e-4, 'softplus reverse')
    # gradients of sum(z * a) + sum(ldj * b) through the backward kernels
    a, b = torch.from_numpy(g['act_a']).to(dev), torch.from_numpy(g['act_b']).to(dev)
    for lay, key in ((L.Sigmoid(temperature=2.5).to(dev), 'sig_dx_2.5'), (L.Softplus(), 'sp_dx')):
        xg = x.clone().requires_grad_(True)
        z, ldj = lay(xg)
        ((z * a).sum() + (ldj * b).sum()).backward()
        assert_close(xg.grad.cpu().numpy(), g[key], 1e-5, 1e-5, key)


@pytest.mark.parametrize('shape', [(3, 1), (1025, 33), (4, 5, 64)])
def test_activation_layers_match_oracle(shape):
    x = synth.uniform('actx', shape) * 12.0
    for T in (1.0, 0.7):
        z, ldj = L.Sigmoid(temperature=T).to(dev)(x.to(dev))
        zo, lo = O.sigmoid_layer(x, T)
        assert_close(z.cpu().numpy(), zo.numpy(), 1e-6, 1e-6, 'sigmoid z'); assert_close(ldj.cpu().numpy(), lo.numpy(), 1e-5, 1e-4, 'sigmoid ldj')
    z, ldj = L.Softplus()(x.to(dev))
    zo, lo = O.softplus_layer(x)
    assert_close(z.cpu().numpy(), zo.numpy(), 1e-6, 1e-6, 'softplus z'); assert_close(ldj.cpu().numpy(), lo.numpy(), 1e-5, 1e-4, 'softplus ldj')
    ok = zo > 1e-2                                         # log(1 - exp(-z)) amplifies one ulp of exp(-z) by 1 / z: compare on the same z, away from 0
    assert_close(L.Softplus().reverse(zo.to(dev)).cpu()[ok].numpy(), O.softplus_layer_reverse(zo)[ok].numpy(), 1e-4, 1e-4, 'softplus reverse')


def test_student_mixture_matches_reference_fixture():
    from tests.golden.make_golden_extras import STUDENT
    g = _extras()
    dist = L.StudentMixtureDistribution(STUDENT['size'], mixtures=STUDENT['mixtures'])
    sd = dist.state_dict(); synth.fill_state(sd, 'student'); dist.load_state_dict(sd)
    dist = dist.to(dev)
    with torch.no_grad():
        logp = dist.log_prob(torch.from_numpy(g['stu_x']).to(dev))
    assert_close(logp.cpu().numpy(), g['stu_logp64'].astype('float32'), 1e-5, 1e-3, 'student log_prob vs the float64 reference')


@pytest.mark.parametrize('B,M,size', [(9, 2, (8, 16, 16)), (300, 10, (4, 3, 2)), (1, 1, (1, 1, 1))])
def test_student_mixture_matches_oracle(B, M, size):
    dist = L.StudentMixtureDistribution(size, mixtures=M)
    sd = dist.state_dict(); synth.fill_state(sd, f'stu{B}'); dist.load_state_dict(sd)
    x = synth.uniform('stux', (B,) + tuple(size)) * 2.0
    ref = O.student_mixture_log_prob({k: v.double() for k, v in dist.state_dict().items() if k in ('mG', 'sG', 'wG', 'mS', 'sS', 'wS', 'vS')}, x.double())
    with torch.no_grad():
        logp = dist.to(dev).log_prob(x.to(dev))
    assert_close(logp.cpu().numpy(), ref.float().numpy(), 2e-5, 1e-3, 'student log_prob')
