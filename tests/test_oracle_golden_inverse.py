"""Pin the oracle's inverse direction and score epilogue against outputs of the unmodified reference
(tests/golden/inv_*.npz, score_*.npz from make_golden_inverse.py).  CPU only."""
import os
import numpy as np
import pytest
import torch

from contextflow_b200 import synth
from oracle import flow_oracle as O
from tests.golden.cases import CASES, INVERSE_CHAIN, INVERSE_COUPLING, SCORE_CASES
from tests.helpers import GOLD, assert_close, case_inputs, golden_state, load_golden


def load_inv(name):
    return dict(np.load(os.path.join(GOLD, f'inv_{name}.npz'), allow_pickle=False))


def floor_agrees(got, prefloor_ref, ref, what):
    """floor() is discontinuous: values within 1e-3 of an integer may land on either side."""
    frac = prefloor_ref - np.floor(prefloor_ref)
    safe = (frac > 1e-3) & (frac < 1 - 1e-3)
    assert np.array_equal(np.asarray(got)[safe], ref[safe]), f'{what}: floor mismatch away from integer boundaries'
    assert np.abs(np.asarray(got) - ref).max() <= 1.0, what


@pytest.mark.parametrize('name', INVERSE_CHAIN)
def test_oracle_reverse_chain_matches_reference(name):
    case = CASES[name]
    g, gi = load_golden(name), load_inv(name)
    stack, state = golden_state(g, case)
    _, ctx = case_inputs(case)
    seen = {}
    x = O.reverse(stack, state, torch.from_numpy(g['z']), ctx, synth.NoiseTape('unused'), torch.float32,
                  trace=lambda lay, v: seen.__setitem__(int(lay['key']), v.clone()))
    for i, v in seen.items():
        if stack['layers'][i]['op'] == 'dequant':
            continue
        ref = gi[f'rsum_{i}']
        vd = v.double()
        assert_close(np.array([vd.sum().item(), vd.abs().sum().item()]), ref, 0.0, 1e-5 * float(ref[1]) + 1e-6, f'{name} reverse layer {i}')
    if 'x_prefloor' in gi:
        assert_close(seen[1].numpy(), gi['x_prefloor'], 1e-5, 1e-4, f'{name} value before the floor')
        floor_agrees(x.numpy(), gi['x_prefloor'], gi['x_rec'], name)
    else:
        assert_close(x.numpy(), gi['x_rec'], 1e-5, 1e-6, f'{name} x_rec')


@pytest.mark.parametrize('name', INVERSE_CHAIN)
def test_reference_reverse_inverts_its_forward(name):
    """Sanity of the fixture itself: the reference's reverse chain returns the golden input (images: exactly, after the floor)."""
    case = CASES[name]
    x, _ = case_inputs(case)
    gi = load_inv(name)
    if 'x_prefloor' in gi:
        assert np.array_equal(gi['x_rec'], x.numpy())
    else:
        assert_close(gi['x_rec'], x.numpy(), 1e-4, 1e-5, name)


@pytest.mark.parametrize('name', INVERSE_COUPLING)
def test_oracle_coupling_reverse_with_context_matches_reference(name):
    case = CASES[name]
    g, gi = load_golden(name), load_inv(name)
    stack, state = golden_state(g, case)
    _, ctx = case_inputs(case)
    shapes = coupling_output_shapes(stack, case)
    for i in gi['layers'].tolist():
        lay = stack['layers'][i]
        zin = synth.NoiseTape(f'invin{i}').randn(shapes[i])
        xr = O.reverse_layer(O._P(state, torch.float32), state, lay, zin, ctx, synth.NoiseTape(f'invnoise{i}'), torch.float32)
        ref = gi[f'rsum_{i}']
        xd = xr.double()
        assert_close(np.array([xd.sum().item(), xd.abs().sum().item()]), ref, 0.0, 1e-5 * float(ref[1]) + 1e-6, f'{name} coupling {i} reverse sum')
        if f'x_{i}' in gi:
            assert_close(xr.numpy(), gi[f'x_{i}'], 1e-4, 1e-5, f'{name} coupling {i} reverse')


def coupling_output_shapes(stack, case):
    """Output shape of every layer, by walking the descriptors (no arithmetic)."""
    C, H, W = case['conf']['data_size']
    B = case['B']
    out = {}
    for i, lay in enumerate(stack['layers']):
        op = lay['op']
        if op == 'augment':
            C += lay['size'][0]
        elif op == 'squeeze':
            p1, p2 = lay['p']; C, H, W = C * p1 * p2, H // p1, W // p2
        elif op == 'permute':
            C, H = H, C
        elif op == 'splitprior':
            C //= 2
        out[i] = (B, C, H, W)
    return out


@pytest.mark.parametrize('name', sorted(SCORE_CASES))
def test_oracle_score_epilogue_matches_reference(name):
    g = dict(np.load(os.path.join(GOLD, f'score_{name}.npz'), allow_pickle=False))
    w = torch.from_numpy(g['weight']) if 'weight' in g else None
    out = O.score_epilogue(torch.from_numpy(g['logp']), float(g['dim_inv']), torch.from_numpy(g['gt']), w)
    B, M = g['logp'].shape
    assert_close(out['scaled'].numpy(), g['scaled'], 1e-6, 1e-7, 'scaled')
    assert_close(out['lse'].numpy(), g['lse'], 1e-6, 1e-6, 'lse')
    assert np.array_equal(out['argmax'].numpy(), g['argmax'])
    assert_close(out['last'].numpy(), g['last'], 1e-6, 1e-7, 'last')
    if 'softmax1' in g:
        assert_close(out['softmax1'].numpy(), g['softmax1'], 1e-5, 1e-7, 'softmax1')
    s = out['sums'].double()
    assert_close(s[0] / B, g['uns_crit'], 1e-5, 1e-6, 'cost_uns (criterion)')
    assert_close(s[1] / (B * M), g['uns_none'], 1e-5, 1e-6, 'cost_uns (no criterion)')
    assert_close(s[2] / s[3], g['sup'], 1e-5, 1e-6, 'cost_sup')
