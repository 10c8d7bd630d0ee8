"""GPU parity: the CUDA path (through the C ABI) against the committed reference goldens and against the CPU oracle.
Gates: z rtol 1e-4 / atol 1e-5; ldj, log-prob rtol 1e-4 / atol 1e-3; bits-per-dim 1e-3; index work bit exact."""
import numpy as np
import pytest
import torch

from contextflow_b200 import builder, rng, synth
from oracle import flow_oracle as O
from tests.golden.cases import CASES
from tests.helpers import (BPD_ATOL, BPD_RTOL, L_ATOL, L_RTOL, Z_ATOL, Z_RTOL, assert_close, case_inputs, golden_state, load_golden)

pytestmark = pytest.mark.gpu


def build_cuda_model(case):
    conf = case['conf']
    model = builder.build_named(conf)
    sd = model.state_dict()
    synth.fill_state(sd, case.get('wseed', 'w0'))
    if case.get('fresh_actnorm'):
        for k in sd:
            if k.endswith('.initialized'):
                sd[k].fill_(0)
    model.load_state_dict(sd)
    return model.cuda().eval()


def run_traced(model, x, ctx, tape):
    rec = {}
    hooks = []
    for i, m in enumerate(model.sequence_modules):
        hooks.append(m.register_forward_hook(
            lambda mod, inp, out, i=i: rec.__setitem__(i, (out[0].detach().cpu(), out[1].detach().cpu(), inp[0].detach().cpu()))))
    with torch.no_grad(), rng.use_source(tape):
        z, logp = model(x.cuda(), ctx.cuda())
    for h in hooks:
        h.remove()
    torch.cuda.synchronize()
    return z.cpu(), logp.cpu(), rec


@pytest.mark.parametrize('name', sorted(CASES))
def test_cuda_matches_reference_golden(name):
    case = CASES[name]
    g = load_golden(name)
    model = build_cuda_model(case)
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == g['keys']
    x, ctx = case_inputs(case)
    tape = synth.NoiseTape(case.get('nseed', 'noise0'))
    z, logp, rec = run_traced(model, x, ctx, tape)
    assert [tuple(d) for d in tape.log] == [(k, tuple(s)) for k, s in g['draws']], 'RNG contract: draw order / shapes'
    for i in range(int(g['n_layers'])):
        zi, ldj = rec[i][:2]
        assert_close(ldj.numpy(), g[f'ldj_{i}'], L_RTOL, L_ATOL, f'{name} layer {i} {g["layer_types"][i]} ldj')
        zd = zi.double()
        ref = g[f'zsum_{i}']
        assert_close(np.array([zd.sum().item(), zd.abs().sum().item()]), ref, 0.0, Z_RTOL * float(ref[1]) + 1e-5,
                     f'{name} layer {i} {g["layer_types"][i]} z-sum')
    assert_close(z.numpy(), g['z'], Z_RTOL, Z_ATOL * max(1.0, float(np.abs(g['z']).max())), f'{name} z')
    assert_close(logp.numpy(), g['logp'], L_RTOL, L_ATOL, f'{name} logp')
    ds = case['conf']['data_size']
    assert_close(O.bits_per_dim(logp, ds).numpy(), O.bits_per_dim(torch.from_numpy(g['logp']), ds).numpy(), BPD_RTOL, BPD_ATOL, f'{name} bpd')
    if case.get('fresh_actnorm'):
        sd = model.state_dict()
        for k, v in g.items():
            if k.startswith('post:'):
                assert_close(sd[k[5:]].cpu().numpy(), v, 1e-4, 1e-5, f'{name} {k}')


@pytest.mark.parametrize('name,B', [('cfg1', 9), ('cfg2', 11), ('cfg3', 5), ('cfg4', 67), ('cifar_conventional', 6), ('msl_conv', 19),
                                    ('mnist_maf', 7), ('msl_maf', 33), ('cifar_gen', 5), ('smd_trans', 21)])
def test_cuda_matches_oracle_per_layer(name, B):
    """Fresh seeded inputs (not in the goldens), ragged batch sizes; every layer's full z and ldj against the oracle."""
    case = dict(CASES[name], B=B, iseed='in1', nseed='noise1')
    g = load_golden(name)
    stack, state = golden_state(g, case)
    model = build_cuda_model(case)
    x, ctx = case_inputs(case)
    ora = {}
    O.forward(stack, state, x, ctx, synth.NoiseTape('noise1'), torch.float32, lambda lay, z, ldj: ora.__setitem__(int(lay['key']), (z.clone(), ldj.clone())))
    _, logp_o = O.forward(stack, state, x, ctx, synth.NoiseTape('noise1'), torch.float32)
    z, logp, rec = run_traced(model, x, ctx, synth.NoiseTape('noise1'))
    for i in sorted(ora):
        zo, lo = ora[i]
        scale = max(1.0, float(zo.abs().max()))
        op = stack['layers'][i]['op']
        if op in ('squeeze', 'permute'):        # index work: bit exact on the layer's own input
            want = O.squeeze(rec[i][2], stack['layers'][i]['p']) if op == 'squeeze' else O.permute_chw(rec[i][2])
            assert torch.equal(rec[i][0], want), f'{name} layer {i}: index op must be bit exact'
        if True:
            assert_close(rec[i][0].numpy(), zo.numpy(), Z_RTOL, Z_ATOL * scale, f'{name} layer {i} {stack["layers"][i]["op"]} z')
        assert_close(rec[i][1].numpy(), lo.numpy(), L_RTOL, L_ATOL, f'{name} layer {i} {stack["layers"][i]["op"]} ldj')
    assert_close(logp.numpy(), logp_o.numpy(), L_RTOL, L_ATOL, f'{name} logp')


def test_empty_batch_and_single_sample():
    case = CASES['cfg4']
    model = build_cuda_model(case)
    x, ctx = synth.make_inputs(case['conf'], 1, 'in2')
    with torch.no_grad(), rng.use_source(synth.NoiseTape('n2')):
        z, lp = model(x.cuda(), ctx.cuda())
    assert lp.shape == (1, 1) and torch.isfinite(lp).all()
    with torch.no_grad(), rng.use_source(synth.NoiseTape('n2')):
        z0, lp0 = model(x[:0].cuda(), ctx[:0].cuda())
    assert lp0.shape == (0, 1)


def test_cpu_tensor_is_rejected_loudly():
    from contextflow_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.squeeze(torch.zeros(1, 1, 2, 2), 2, 2)


# ---------------------------------------------------------------------------------------------------- fused log_prob plan
FUSED_EXPECTED = {'cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg1_init', 'cfg4_init', 'cfg2_init', 'mnist_onehot_uniform', 'mnist_eye_uniform',
                  'mnist_eye_vardeq2', 'atm_argmax2', 'cifar_conventional', 'smap_conventional', 'msl_conv', 'mnist28'}


@pytest.mark.parametrize('name', sorted(CASES))
def test_fused_log_prob_matches_reference_golden(name):
    """`log_prob` through the fused plan (layers/_fastpath.py: one-kernel prologue, context pre-pass, Conv1x1+ActNorm kernel, single
    ldj reduction) against the reference's log-probabilities, and against the layer-by-layer CUDA forward on the same draws."""
    case = CASES[name]
    g = load_golden(name)
    model = build_cuda_model(case)
    x, ctx = case_inputs(case)
    xc, cc = x.cuda(), ctx.cuda()
    with torch.no_grad():
        if case.get('fresh_actnorm'):                            # the first batch initialises ActNorm layer by layer (actnorm.py:46,53)
            with rng.use_source(synth.NoiseTape(case.get('nseed', 'noise0'))):
                model(xc, cc)
        tape = synth.NoiseTape(case.get('nseed', 'noise0'))
        with rng.use_source(tape):
            logp = model.log_prob(xc, cc)
        with rng.use_source(synth.NoiseTape(case.get('nseed', 'noise0'))):
            logp_layers = model(xc, cc)[1]
    with torch.no_grad():
        fp_ok = model._fastpath.usable(xc, cc)
    assert fp_ok or name not in FUSED_EXPECTED, f'{name}: the fused plan should cover this stack'
    assert [tuple(d) for d in tape.log] == [(k, tuple(s)) for k, s in g['draws']], 'RNG contract: draw order / shapes'
    assert_close(logp.cpu().numpy(), g['logp'], L_RTOL, L_ATOL, f'{name} fused logp vs reference')
    assert_close(logp.cpu().numpy(), logp_layers.cpu().numpy(), 1e-5, 1e-3, f'{name} fused vs layer-by-layer')
    ds = case['conf']['data_size']
    assert_close(O.bits_per_dim(logp.cpu(), ds).numpy(), O.bits_per_dim(torch.from_numpy(g['logp']), ds).numpy(), BPD_RTOL, BPD_ATOL, f'{name} bpd')


def test_fused_plan_matches_oracle_ragged_batch():
    case = dict(CASES['cfg2'], B=37, iseed='in3', nseed='noise3')
    g = load_golden('cfg2')
    stack, state = golden_state(g, case)
    model = build_cuda_model(case)
    x, ctx = case_inputs(case)
    _, want = O.forward(stack, state, x, ctx, synth.NoiseTape('noise3'), torch.float32)
    with torch.no_grad(), rng.use_source(synth.NoiseTape('noise3')):
        got = model.log_prob(x.cuda(), ctx.cuda())
        assert model._fastpath.usable(x.cuda(), ctx.cuda())
    assert_close(got.cpu().numpy(), want.numpy(), L_RTOL, L_ATOL, 'cfg2 fused logp vs oracle, B=37')


def test_cuda_graph_replay_matches_eager():
    """The thing bench.py times is the CUDA-graph replay: with the same generator seed before the eager call and before the replay the
    Philox draws are identical, so the log-probs must be identical BIT FOR BIT (same kernels, same inputs, same noise)."""
    case = CASES['cfg2']
    model = build_cuda_model(case)
    x, ctx = synth.make_inputs(case['conf'], 16, 'in4')
    xc, cc = x.cuda(), ctx.cuda()
    with torch.no_grad():
        model.log_prob(xc, cc)                                    # ActNorm initialisation, weight packing
        torch.manual_seed(5); a = model.log_prob(xc, cc)
        model.enable_cuda_graphs()
        model.log_prob(xc, cc)                                    # captures (its warm-up consumes draws)
        torch.manual_seed(5); b = model.log_prob(xc, cc)
        b2 = model.log_prob(xc, cc)
        torch.manual_seed(5); b3 = model.log_prob(xc, cc)
    assert a.shape == b.shape == (16, case['conf']['mixtures']) and torch.isfinite(b).all() and torch.isfinite(b2).all()
    assert torch.equal(a, b), f'graph replay differs from the eager launch sequence: max abs diff {(a - b).abs().max().item():.3e}'
    assert torch.equal(b, b3)
    assert not torch.equal(b, b2)                                 # replays advance the Philox offset: fresh noise every batch


# ---- parity at the batch sizes bench.py measures (VERDICT r1 "what's weak" 1-3) ------------------------------------------------------
def _bench_like_inputs(conf, B, seed):
    gen = torch.Generator().manual_seed(seed)
    C, H, W = conf['data_size']
    x = torch.randint(0, 256, (B, C, H, W), generator=gen).float() if conf['image'] else torch.rand(B, C, H, W, generator=gen)
    ctx = torch.stack([torch.randint(0, k, (B,), generator=gen) for k in conf['contexts']], 1)
    return x, ctx


@pytest.mark.parametrize('name,B', [('cfg2', 8192), ('cfg4', 131072), ('cfg1', 8192), ('cfg3', 1024)])
def test_bench_batch_matches_oracle_and_graph_replay(name, B):
    """The configuration bench.py times: a full default batch through (1) the eager fused plan with recorded draws, (2) the CUDA-graph
    replay under the same seed -- bit-identical to (1) -- and (3) the CPU oracle on a 256-row subsample replaying the recorded noise rows.
    At these sizes the context-bucketed mixture kernel (B >= 32 x context tuples), multi-tile persistent CTAs and the single batched
    encoder-noise draw are what runs; none of them is reached by the small-batch golden cases."""
    from oracle import check
    case = CASES[name]
    conf = case['conf']
    model = build_cuda_model(case)
    x, ctx = _bench_like_inputs(conf, B, 77)
    xc, cc = x.cuda(), ctx.cuda()
    with torch.no_grad():
        model.log_prob(xc[:64], cc[:64])                          # packing / lazy state
    logp, rec = check.recorded_log_prob(model, xc, cc, seed=11)
    with torch.no_grad():
        assert model._fastpath.usable(xc, cc)
    res = check.rows_parity(model, conf, xc, cc, seed=11, n_rows=256, logp=logp, rec=rec)
    assert res['ok'], f'{name} B={B}: {res}'
    model.enable_cuda_graphs()
    with torch.no_grad():
        model.log_prob(xc, cc)                                    # capture
        torch.manual_seed(11); rep = model.log_prob(xc, cc)
    assert torch.equal(rep, logp), f'{name} B={B}: graph replay differs from eager, max abs {(rep - logp).abs().max().item():.3e}'


def _encoder_outputs(model, ctx, tape):
    """{layer index: (c, logp_c)} of every flow layer's context encoder, evaluated by the batched encoder launch the fused plan uses."""
    out = {}
    with torch.no_grad(), rng.use_source(tape):
        groups = model._encoder_groups()
        assert len(groups) == 1
        batch = next(iter(groups.values()))
        batch.run(ctx)
        plans = {id(p): (p, f) for p, f in batch.members}
        for i, m in enumerate(model.sequence_modules):
            p = getattr(m, '_plan', None)
            if p is not None and id(p) in plans:
                out[i] = (p.preset[0].clone(), p.preset[1].clone()); p.preset = None
    return out


@pytest.mark.parametrize('name', ['cfg3', 'cfg2', 'atm_argmax2', 'mnist_eye_vardeq2', 'mnist_onehot_uniform'])
def test_every_context_value_is_encoded_exactly(name):
    """Exhaustive sweep of the context space (all 68 ATM entities, all 15 x 5 CIFAR corruption tuples, ...): the integer part of every
    encoder's output -- argmax sign bits MSB first with the zero pad column (dequantize.py:196-211,239-268), the one-hot / eye code under
    the dequantisation noise (dequantize.py:55-63,107-116) -- is BIT EXACT against the oracle for every context value and every layer;
    the full (c, logp_c) agrees within the z / ldj gates."""
    case = CASES[name]
    conf = case['conf']
    model = build_cuda_model(case)
    cards = conf['contexts']
    grids = torch.cartesian_prod(*[torch.arange(k) for k in cards]).reshape(-1, len(cards))
    ctx = torch.cat([grids, grids.flip(0)], 0)                    # every tuple twice, different noise rows
    B = ctx.shape[0]
    assert B == 2 * int(np.prod(cards))
    g = load_golden(name)
    stack, state = golden_state(g, case)
    got = _encoder_outputs(model, ctx.cuda(), synth.NoiseTape('sweep'))
    assert len(got) >= 3
    tape = synth.NoiseTape('sweep')
    P = O._P(state, torch.float32)
    n_checked = 0
    for lay in stack['layers']:
        if lay.get('enc') is None or lay['op'] == 'splitprior':
            continue
        i = int(lay['key'])
        spec = lay['enc']
        c_ref, lp_ref = O.context_encode(P, state, f'{i}.context_net', spec, ctx, tape, torch.float32)
        c, lp = got[i][0].cpu(), got[i][1].cpu()
        assert_close(c.numpy(), c_ref.numpy(), Z_RTOL, Z_ATOL, f'{name} layer {i} encoder c')
        assert_close(lp.numpy(), lp_ref.numpy(), L_RTOL, L_ATOL, f'{name} layer {i} encoder logp_c')
        if spec['type'] == 'argmax':
            bits = torch.cat([O.int_to_bits(ctx[:, j], b) for j, b in enumerate(spec['bits'])], -1)
            if bits.shape[-1] % 2:
                bits = torch.cat([bits, torch.zeros(B, 1, dtype=bits.dtype)], -1)
            assert torch.equal(torch.sign(c), (bits * 2 - 1).float()), f'{name} layer {i}: argmax sign bits'
        else:                                                     # uniform / vardeq: c * qbins = code + noise in [0, 1)
            q = state[f'{i}.context_net.1.qbins']
            code = torch.cat([torch.nn.functional.one_hot(ctx[:, j], k) for j, k in enumerate(cards)], 1) if spec['emb'] == 'onehot' else ctx
            assert torch.equal(torch.floor(c * q).long(), code.long()), f'{name} layer {i}: integer code under the dequantisation noise'
        n_checked += 1
    assert n_checked == len(got)
