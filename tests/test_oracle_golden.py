"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

from contextflow_b200 import synth
from oracle import flow_oracle as O
from tests.golden.cases import CASES
from tests.helpers import load_golden, golden_state, case_inputs, assert_close


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_reference_golden(name):
    case = CASES[name]
    g = load_golden(name)
    stack, state = golden_state(g, case)
    assert len(stack['layers']) == int(g['n_layers'])
    x, ctx = case_inputs(case)
    tape = synth.NoiseTape(case.get('nseed', 'noise0'))
    seen = []

    def trace(lay, z, ldj):
        i = int(lay['key'])
        assert_close(ldj.numpy(), g[f'ldj_{i}'], 1e-5, 1e-4, f'{name} layer {i} {lay["op"]} ldj')
        zd = z.double()
        assert_close(np.array([zd.sum().item(), zd.abs().sum().item()]), g[f'zsum_{i}'], 0.0, 2e-6 * float(g[f'zsum_{i}'][1]) + 1e-6, f'{name} layer {i} z-sum')
        seen.append(i)

    z, logp = O.forward(stack, state, x, ctx, tape, torch.float32, trace)
    assert seen == list(range(int(g['n_layers'])))
    # the RNG contract: same draws, same order, same shapes as the reference made
    assert [tuple(d) for d in tape.log] == [(k, tuple(s)) for k, s in g['draws']]
    assert_close(z.numpy(), g['z'], 1e-5, 1e-5, f'{name} z')
    assert_close(logp.numpy(), g['logp'], 1e-5, 1e-4, f'{name} logp')
    if case.get('fresh_actnorm'):
        for k, v in g.items():
            if k.startswith('post:'):
                assert_close(state[k[5:]].numpy(), v, 1e-5, 1e-6, f'{name} {k}')


def test_oracle_float64_agrees_with_float32():
    case = CASES['cfg2']
    g = load_golden('cfg2')
    stack, state = golden_state(g, case)
    x, ctx = case_inputs(case)
    _, lp64 = O.forward(stack, state, x, ctx, synth.NoiseTape('noise0'), torch.float64)
    assert_close(lp64.numpy(), g['logp'], 2e-5, 1e-3, 'fp64 oracle vs fp32 reference')


def test_index_ops_bit_exact():
    x = synth.uniform('sq', (2, 3, 4, 6))
    y = O.squeeze(x, (2, 2))
    for c in range(3):
        for i in range(2):
            for j in range(2):
                assert torch.equal(y[:, c * 4 + i * 2 + j], x[:, c, i::2, j::2])
    bits = O.int_to_bits(torch.tensor([0, 1, 5, 67]), 7)
    assert bits.tolist() == [[0] * 7, [0, 0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 1, 0, 1], [1, 0, 0, 0, 0, 1, 1]]
