"""GPU parity of the inverse (sampling) direction and of the loss / score epilogue, through the C ABI:
against fixtures made by the unmodified reference (tests/golden/inv_*.npz, score_*.npz), against the oracle on fresh inputs,
and through round trips forward -> reverse at full batch sizes."""
import os
import numpy as np
import pytest
import torch

from contextflow_b200 import builder, ops, rng, synth
from oracle import flow_oracle as O
from tests.golden.cases import CASES, INVERSE_CHAIN, INVERSE_COUPLING, SCORE_CASES
from tests.helpers import GOLD, Z_ATOL, Z_RTOL, assert_close, case_inputs, golden_state, load_golden
from tests.test_oracle_golden_inverse import coupling_output_shapes, floor_agrees, load_inv

pytestmark = pytest.mark.gpu


def build_cuda_model(case):
    conf = case['conf']
    model = builder.build_named(conf)
    sd = model.state_dict(); synth.fill_state(sd, case.get('wseed', 'w0')); model.load_state_dict(sd)
    return model.cuda().eval()


@pytest.mark.parametrize('name', INVERSE_CHAIN)
def test_reverse_chain_matches_reference_golden(name):
    case = CASES[name]
    g, gi = load_golden(name), load_inv(name)
    model = build_cuda_model(case)
    _, ctx = case_inputs(case)
    z = torch.from_numpy(g['z']).cuda(); ctx = ctx.cuda()
    mods = list(model.sequence_modules)
    out, pre = z, None
    with torch.no_grad():
        for i in reversed(range(len(mods))):                         # layer by layer, as flowsequential.py:34-37
            if type(mods[i]).__name__ == 'Dequantization':
                pre = out.cpu().numpy()
            out = mods[i].reverse(out, ctx)
            if type(mods[i]).__name__ != 'Dequantization':
                ref = gi[f'rsum_{i}']
                od = out.double()
                assert_close(np.array([od.sum().item(), od.abs().sum().item()]), ref, 0.0, Z_RTOL * float(ref[1]) + 1e-5,
                             f'{name} reverse layer {i} {type(mods[i]).__name__}')
        fused = model.reverse(z, ctx)                                 # the same chain with the one-kernel inverse prologue
    if pre is not None:
        assert_close(pre, gi['x_prefloor'], Z_RTOL, 1e-3, f'{name} value before the floor')
        floor_agrees(out.cpu().numpy(), gi['x_prefloor'], gi['x_rec'], f'{name} layer chain')
        floor_agrees(fused.cpu().numpy(), gi['x_prefloor'], gi['x_rec'], f'{name} fused chain')
    else:
        assert_close(out.cpu().numpy(), gi['x_rec'], Z_RTOL, Z_ATOL, f'{name} x_rec')
        assert torch.equal(out, fused)


@pytest.mark.parametrize('name', INVERSE_COUPLING)
def test_coupling_reverse_with_context_matches_reference_golden(name):
    case = CASES[name]
    g, gi = load_golden(name), load_inv(name)
    model = build_cuda_model(case)
    stack, _ = golden_state(g, case)
    _, ctx = case_inputs(case)
    shapes = coupling_output_shapes(stack, case)
    mods = list(model.sequence_modules)
    for i in gi['layers'].tolist():
        zin = synth.NoiseTape(f'invin{i}').randn(shapes[i]).cuda()
        with torch.no_grad(), rng.use_source(synth.NoiseTape(f'invnoise{i}')):
            xr = mods[i].reverse(zin, ctx.cuda())
        ref = gi[f'rsum_{i}']
        xd = xr.double()
        assert_close(np.array([xd.sum().item(), xd.abs().sum().item()]), ref, 0.0, Z_RTOL * float(ref[1]) + 1e-5, f'{name} coupling {i} sum')
        if f'x_{i}' in gi:
            assert_close(xr.cpu().numpy(), gi[f'x_{i}'], Z_RTOL, Z_ATOL * max(1.0, float(np.abs(gi[f'x_{i}']).max())), f'{name} coupling {i}')
        assert torch.equal(xr[:, : xr.shape[1] // 2], zin[:, : xr.shape[1] // 2]), 'pass-through half is bit exact'


@pytest.mark.parametrize('name,B', [('cfg1', 13), ('cfg4', 70), ('mnist28', 5)])
def test_reverse_chain_matches_oracle_fresh_inputs(name, B):
    case = dict(CASES[name], B=B, iseed='in2')
    g = load_golden(name)
    stack, state = golden_state(g, case)
    model = build_cuda_model(case)
    x, ctx = case_inputs(case)
    out_size = stack['out_size']
    z = synth.NoiseTape('zlat').randn((B,) + tuple(out_size))
    image = stack['layers'][0]['op'] == 'dequant'
    ref = O.reverse(stack, state, z, ctx, synth.NoiseTape('u'), torch.float32, stop_before=1 if image else 0)
    with torch.no_grad():
        if image:
            got = z.cuda()
            for m in reversed(list(model.sequence_modules)[1:]):
                got = m.reverse(got, ctx.cuda())
            assert_close(got.cpu().numpy(), ref.numpy(), Z_RTOL, 1e-3, f'{name} before the floor')
            floor_agrees(model.reverse(z.cuda(), ctx.cuda()).cpu().numpy(), ref.numpy(), ref.floor().numpy(), name)
        else:
            got = model.reverse(z.cuda(), ctx.cuda())
            assert_close(got.cpu().numpy(), ref.numpy(), Z_RTOL, Z_ATOL * max(1.0, float(ref.abs().max())), name)


@pytest.mark.parametrize('name,B', [('cfg1', 4096), ('cfg4', 16384)])
def test_forward_reverse_round_trip_full_batch(name, B):
    """Size-independent property at bench batch sizes: reverse(forward(x)) == x (images: exactly after the floor, up to
    boundary cases; time series: to 1e-4)."""
    case = dict(CASES[name], B=B, iseed='in3')
    model = build_cuda_model(case)
    x, ctx = case_inputs(case)
    x, ctx = x.cuda(), ctx.cuda()
    with torch.no_grad():
        z, _ = model(x, ctx)
        xr = model.reverse(z, ctx)
    assert xr.shape == x.shape
    if name == 'cfg1':
        wrong = (xr != x).float().mean().item()
        assert wrong < 1e-3, f'{wrong:.2e} of the pixels differ after the round trip'
        assert (xr - x).abs().max().item() <= 1.0
    else:
        assert_close(xr.cpu().numpy(), x.cpu().numpy(), 1e-4, 1e-4, name)


@pytest.mark.parametrize('D', [1, 2, 8, 26, 32, 76, 128])
def test_mat_inverse(D):
    A = synth.normal(f'inv{D}', (D, D)) / np.sqrt(D) + torch.eye(D)
    if D >= 2:
        A[0, 0] = 0.0                                                   # forces a row exchange at the first pivot
    inv, flag = ops.mat_inverse(A.cuda())
    assert int(flag.item()) == 0
    ref = torch.linalg.inv(A.double())
    assert_close(inv.cpu().numpy(), ref.numpy(), 1e-5, 1e-6 * float(ref.abs().max()), f'inverse D={D}')
    assert_close((inv.cpu().double() @ A.double()).numpy(), np.eye(D), 0.0, 1e-5, 'A^-1 A = I')


def test_mat_inverse_flags_singular():
    A = torch.ones(4, 4)
    _, flag = ops.mat_inverse(A.cuda())
    assert int(flag.item()) == 1


def test_gmm_sample_given_draws_matches_oracle():
    case = CASES['cfg1']
    g = load_golden('cfg1')
    stack, state = golden_state(g, case)
    model = build_cuda_model(case)
    B = 257
    size = tuple(stack['out_size'])
    comp = (synth.NoiseTape('comp').rand((B,)) * 8).long().clamp(max=7)
    eps = synth.NoiseTape('eps').randn((B,) + size)
    ref = O.gmm_sample_given(O._P(state, torch.float32), stack['base'], comp, eps)
    got = ops.gmm_sample(model.dist.mG.detach(), model.dist.sG.detach(), comp.cuda(), eps.cuda(), m=1)
    assert_close(got.cpu().numpy(), ref.numpy(), 1e-5, 1e-6, 'gmm sample')


def test_sample_statistics_and_errors():
    """FlowSequential.sample (flowsequential.py:32-39): shape / range for the image generalist; the component frequencies of the base
    draw follow softmax(wG[1]); M = 1 and split-prior models fail the way the reference does."""
    model = build_cuda_model(CASES['cfg1'])
    with torch.no_grad():
        xs = model.sample(64)
        assert xs.shape == (64, 1, 32, 32) and torch.isfinite(xs).all()
        assert torch.equal(xs, xs.floor()) and xs.min().item() >= -1 and xs.max().item() <= 256
        n = 40000
        torch.manual_seed(5)
        z, logp = model.dist.sample(n)
        assert z.shape == (n,) + tuple(model.dist.size) and logp.shape == (n, model.dist.M)
        # nearest component mean of mixture 1 identifies the draw (means are far apart relative to scale at this fill)
        w = torch.softmax(model.dist.wG[1].double(), -1).cpu()
        mean = (torch.softmax(model.dist.wG[1], -1)[:, None] * model.dist.mG[1].flatten(1)).sum(0)
        got = z.flatten(1).mean(0)
        sd = torch.sqrt(((torch.nn.functional.softplus(model.dist.sG[1]) ** 2 + model.dist.mG[1] ** 2).flatten(1) * torch.softmax(model.dist.wG[1], -1)[:, None]).sum(0) - mean ** 2)
        zscore = ((got - mean) / (sd / np.sqrt(n))).abs()
        assert zscore.max().item() < 6.0, f'sample mean off by {zscore.max().item():.1f} standard errors'
        assert abs(float(w.sum()) - 1.0) < 1e-6
    m1 = build_cuda_model(CASES['cfg4'])
    if m1.dist.M == 1:
        with pytest.raises(IndexError), torch.no_grad():
            m1.sample(4)
    sp = build_cuda_model(dict(conf=synth.variant('cfg2', num_blocks=2, block_size=1)))
    with pytest.raises((AttributeError, NotImplementedError)), torch.no_grad():
        sp.reverse(torch.zeros(2, 32, 8, 8, device='cuda'), torch.zeros(2, 2, dtype=torch.int64, device='cuda'))


@pytest.mark.parametrize('name', sorted(SCORE_CASES))
def test_score_epilogue_matches_reference_golden(name):
    g = dict(np.load(os.path.join(GOLD, f'score_{name}.npz'), allow_pickle=False))
    w = torch.from_numpy(g['weight']).cuda() if 'weight' in g else None
    out = ops.score_epilogue(torch.from_numpy(g['logp']).cuda(), float(g['dim_inv']), torch.from_numpy(g['gt']).cuda(), w)
    B, M = g['logp'].shape
    assert_close(out['scaled'].cpu().numpy(), g['scaled'], 1e-6, 1e-7, 'scaled')
    assert_close(out['lse'].cpu().numpy(), g['lse'], 1e-5, 1e-6, 'lse')
    assert np.array_equal(out['argmax'].cpu().numpy(), g['argmax'])
    assert_close(out['last'].cpu().numpy(), g['last'], 1e-6, 1e-7, 'last')
    if 'softmax1' in g:
        assert_close(out['softmax1'].cpu().numpy(), g['softmax1'], 1e-4, 1e-7, 'softmax1')
    s = out['sums'].cpu().double()
    assert_close(s[0] / B, g['uns_crit'], 1e-5, 1e-6, 'cost_uns (criterion)')
    assert_close(s[1] / (B * M), g['uns_none'], 1e-5, 1e-6, 'cost_uns (no criterion)')
    assert_close(s[2] / s[3], g['sup'], 1e-5, 1e-6, 'cost_sup')


@pytest.mark.parametrize('B,M', [(100003, 10), (5001, 40), (1, 2), (255, 1)])
def test_score_epilogue_large_batch_deterministic(B, M):
    logp = synth.NoiseTape('bigscore').randn((B, M)) * 300.0 - 9000.0
    gt = (synth.NoiseTape('biggt').rand((B,)) * M).long().clamp(max=M - 1)
    ref = O.score_epilogue(logp, 1.0 / 3072, gt, None)
    a = ops.score_epilogue(logp.cuda(), 1.0 / 3072, gt.cuda(), None)
    b = ops.score_epilogue(logp.cuda(), 1.0 / 3072, gt.cuda(), None)
    for k in ('scaled', 'lse', 'softmax1', 'last', 'argmax', 'sums'):
        assert torch.equal(a[k], b[k]), f'{k} differs between two runs'
    assert torch.equal(a['scaled'].cpu(), ref['scaled'])
    assert_close(a['softmax1'].cpu().numpy(), ref['softmax1'].numpy(), 1e-4, 1e-7, 'softmax1')
    assert torch.equal(a['argmax'].cpu(), ref['argmax'])
    assert_close(a['lse'].cpu().numpy(), ref['lse'].numpy(), 1e-5, 1e-6, 'lse')
    assert_close(a['sums'].cpu().numpy(), ref['sums'].numpy(), 1e-5, 1e-3, 'sums')


def test_empty_and_single_sample_batches():
    """Edge cases the reference's ops accept: B = 0 (empty tensors flow through) and B = 1, for the inverse / training / window ops."""
    from contextflow_b200.windows import WindowedScorer
    for B in (0, 1):
        z = torch.zeros(B, 8, 4, 4, device='cuda'); h = torch.zeros(B, 8, 4, 4, device='cuda')
        assert ops.coupling_inv(z, h).shape == (B, 8, 4, 4)
        assert ops.actnorm_inv(z, torch.zeros(8, device='cuda'), torch.zeros(8, device='cuda')).shape == (B, 8, 4, 4)
        assert ops.unsqueeze(z, 2, 2).shape == (B, 2, 8, 8)
        assert ops.prologue_inv(z, 7, 1.0, 0.0, 256.0, 0.0).shape == (B, 7, 4, 4)
        dx, dh = ops.coupling_bwd(z, h, z)
        assert dx.shape == z.shape and dh.shape == h.shape
        assert ops.maf_coupling(z, torch.zeros(B, 16, 4, 4, device='cuda'))[1].shape == (B,)
        assert ops.conv2d_fwd(z, 8, torch.zeros(5, 8, 3, 3, device='cuda'), torch.zeros(5, device='cuda'), relu=True).shape == (B, 5, 4, 4)
        dW, db = ops.conv2d_bwd_weight(z, 8, torch.zeros(B, 5, 4, 4, device='cuda'), (5, 8, 3, 3))
        assert not dW.any() and not db.any()
    sc = WindowedScorer(torch.rand(1, 3, dtype=torch.float64), 4)
    assert torch.equal(sc.windows(0, 1).cpu(), torch.from_numpy(np.broadcast_to(sc.ts.cpu().float().numpy().T[None, :, :, None], (1, 3, 4, 1)).copy()))
    with pytest.raises(RuntimeError):
        ops.score_epilogue(torch.zeros(0, 3, device='cuda'), 1.0)          # the reference's .mean() over an empty batch is NaN: refused loudly
    one = ops.score_epilogue(torch.tensor([[-3.0, -1.0]], device='cuda'), 0.5)
    assert one['argmax'].item() == 1 and abs(one['lse'].item() - float(torch.logsumexp(torch.tensor([-1.5, -0.5]), 0))) < 1e-6
