"""Pin the oracle's training direction (loss of experiment_ad.py:204-209 + torch autograd over the oracle's op sequence) against
gradients produced by the unmodified reference (tests/golden/train_*.npz from make_golden_training.py).  CPU only."""
import json, os
import numpy as np
import pytest
import torch

from contextflow_b200 import synth
from oracle import flow_oracle as O
from tests.golden.cases import CASES, TRAINING_CASES
from tests.helpers import GOLD, assert_close, case_inputs, golden_state, load_golden


def load_train(name):
    g = dict(np.load(os.path.join(GOLD, f'train_{name}.npz'), allow_pickle=False))
    g['names'] = json.loads(str(g['names']))
    return g


def labels(name, B, M):
    return (synth.NoiseTape(f'traingt:{name}').rand((B,)) * M).long().clamp(max=M - 1)


def check_grads(name, grads, gt_gold, rtol=2e-4, ref_factor=3.0):
    """grads: {param name: tensor}.  Truth = the reference's own float64 gradients (g64); a gradient passes when it is within
    rtol of the gradient's largest entry, or as close to the truth as the reference's float32 run (g) gets, times 3: every parameter
    gradient is a signed sum over batch and pixels (Conv1x1's dNN and ActNorm's dlogs are differences of large cancelling sums), so
    float32 summation order moves the reference itself by up to ~1e-3 of the largest entry."""
    for k in gt_gold['names']:
        g = grads[k].detach().cpu().double()
        ref_sum = gt_gold[f'gsum:{k}']
        scale = float(ref_sum[2]) + 1e-12
        ref32 = torch.from_numpy(gt_gold[f'g:{k}']).double()
        truth = torch.from_numpy(gt_gold[f'g64:{k}']).double()
        got = g if truth.shape == g.shape else g.flatten()[: truth.numel()]
        err = (got - truth).abs().max().item()
        ref_err = (ref32 - truth).abs().max().item()
        tol = rtol * scale + ref_factor * ref_err + 1e-7
        assert err <= tol, f'{name} grad {k}: max abs err {err:.3e} (float32 reference err {ref_err:.3e}) vs scale {scale:.3e}'
        if truth.shape != g.shape:              # large tensors store their first 512 values: the rest is covered by the checksums
            assert abs(g.sum().item() - ref_sum[0]) <= tol * g.numel() ** 0.5 + rtol * ref_sum[1] + 1e-6, f'{name} grad {k}: sum'
            assert abs(g.abs().sum().item() - ref_sum[1]) <= tol * g.numel() ** 0.5 + rtol * ref_sum[1] + 1e-6, f'{name} grad {k}: abs sum'


@pytest.mark.parametrize('name', sorted(TRAINING_CASES))
def test_oracle_gradients_match_reference(name):
    case, spec = CASES[name], TRAINING_CASES[name]
    g, gt_gold = load_golden(name), load_train(name)
    stack, state = golden_state(g, case)
    for k in gt_gold['names']:
        state[k].requires_grad_(True)
    x, ctx = case_inputs(case)
    _, logp = O.forward(stack, state, x, ctx, synth.NoiseTape(case.get('nseed', 'noise0')), torch.float32)
    w = None if spec['weight'] is None else torch.tensor(spec['weight'])
    cost, sup, uns = O.training_loss(logp, labels(name, case['B'], case['conf']['mixtures']), case['conf']['data_size'], spec['alpha'], spec['criterion'], w)
    assert_close(np.array([cost.item(), sup.item(), uns.item()]), gt_gold['loss'], 1e-5, 1e-6, f'{name} loss')
    cost.backward()
    check_grads(name, {k: state[k].grad for k in gt_gold['names']}, gt_gold, rtol=5e-5)
