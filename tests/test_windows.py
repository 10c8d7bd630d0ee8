"""Sliding-window generation: oracle vs the reference's get_windows fixture (CPU); CUDA kernel vs both, bit exact (GPU)."""
import os
import numpy as np
import pytest
import torch

from contextflow_b200 import synth
from oracle import flow_oracle as O
from tests.helpers import GOLD

G = dict(np.load(os.path.join(GOLD, 'windows.npz')))
NAMES = sorted(k[:-2] for k in G if k.endswith(':x'))


def series(name):
    rows, D, L, stride = (int(v) for v in G[f'{name}:spec'])
    return synth.uniform(f'win:{name}', (rows, D), 0.0, 1.0, torch.float64), L, stride


@pytest.mark.parametrize('name', NAMES)
def test_oracle_windows_match_reference(name):
    ts, L, stride = series(name)
    assert torch.equal(O.sliding_windows(ts.numpy(), L, stride), torch.from_numpy(G[f'{name}:x']))


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_cuda_windows_bit_exact(name, dtype):
    from contextflow_b200 import ops
    from contextflow_b200.windows import WindowedScorer
    ts, L, stride = series(name)
    ref = torch.from_numpy(G[f'{name}:x']) if dtype == torch.float64 else O.sliding_windows(ts.to(dtype).numpy(), L, stride)
    sc = WindowedScorer(ts.to(dtype), L, stride)
    assert len(sc) == ref.shape[0]
    assert torch.equal(sc.windows(0, len(sc)).cpu(), ref)
    assert torch.equal(sc.windows(1, len(sc) - 1).cpu(), ref[1:])
    ends = torch.tensor([0, (len(sc) - 1) * stride, 0], dtype=torch.int64)
    assert torch.equal(ops.windows(ts.to(dtype).cuda(), L, end=ends.cuda()).cpu(), ref[[0, len(sc) - 1, 0]])


@pytest.mark.gpu
def test_windowed_scorer_equals_scoring_materialised_windows():
    """Full SMAP-shaped path at bench size: scores of device-generated windows == scores of the host-materialised windows."""
    from contextflow_b200 import builder
    from contextflow_b200.windows import WindowedScorer
    from tests.golden.cases import CASES
    conf = CASES['cfg4']['conf']
    model = builder.build_named(conf)
    sd = model.state_dict(); synth.fill_state(sd, 'w0'); model.load_state_dict(sd)
    model = model.cuda().eval()
    ts = synth.uniform('win:big', (20000, 25), 0.0, 1.0, torch.float64)
    x_host = O.sliding_windows(ts.numpy(), 8, 1)
    torch.manual_seed(3)
    a = WindowedScorer(ts, 8).scores(model, context_id=7, batch=8192)
    torch.manual_seed(3)
    with torch.no_grad():
        b = torch.cat([model.log_prob(x_host[i:i + 8192].cuda(), context=torch.full((min(8192, 20000 - i), 1), 7, device='cuda')) for i in range(0, 20000, 8192)])
    assert a.shape == (20000, 1) and torch.equal(a, b)
