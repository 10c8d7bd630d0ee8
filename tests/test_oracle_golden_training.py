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


def check_grads(name, grads, gt_gold, rtol=2e-4):
    """grads: {param name: tensor}.  Tolerance: rtol relative to the largest entry of that gradient (sums over the batch reorder)."""
    for k in gt_gold['names']:
        g = grads[k].detach().cpu().double()
        ref_sum = gt_gold[f'gsum:{k}']
        scale = float(ref_sum[2]) + 1e-12
        ref = torch.from_numpy(gt_gold[f'g:{k}']).double()
        got = g if ref.shape == g.shape else g.flatten()[: ref.numel()]
        err = (got - ref).abs().max().item()
        # Every parameter gradient is a signed sum over the batch and the pixels; Conv1x1's dNN is in addition a difference of two large
        # sums (data term and HW * NN^-T) that cancel at a likelihood optimum.  float32 reorderings of those sums move the reference's
        # own result by up to ~1e-3 of the gradient's largest entry (measured against float64 autograd in
        # tests/test_gpu_training.py::test_gradients_match_oracle_autograd_fresh_inputs, which holds the CUDA path to the principled
        # bound: 2e-4 relative or 3x the reference's own float32 error).  Here: rtol for the same-order CPU oracle, 5x for CUDA sums.
        tol = (10 * rtol if k.endswith('.NN') else rtol) * scale + 1e-7
        assert err <= tol, f'{name} grad {k}: max abs err {err:.3e} vs scale {scale:.3e}'
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
