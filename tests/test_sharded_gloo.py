"""N>1 path on CPU: world_size-2 gloo processes shard a global batch, score their slices (the oracle stands in for the
per-rank CUDA forward) and all-gather the log-probs; result must equal the single-process answer row for row."""
import os, sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from contextflow_b200 import synth
    from contextflow_b200.sharded import ShardedLogProb, shard_bounds
    from oracle import flow_oracle as O
    from tests.golden.cases import CASES
    from tests.helpers import golden_state, load_golden
    case = dict(CASES['cfg4'], B=B)
    stack, state = golden_state(load_golden('cfg4'), case)
    x, ctx = synth.make_inputs(case['conf'], B, 'shard')
    eps = synth.normal('shard:eps', (B, 1, 8, 1))          # the Augment draw of cfg4, fixed per global row

    class RowNoise:                                           # every rank draws the rows of ITS slice
        def __init__(self, lo, hi): self.lo, self.hi = lo, hi
        def randn(self, shape): return eps[self.lo:self.hi]
        def rand(self, shape): raise AssertionError
    lo, hi = shard_bounds(B, world, rank)
    local = lambda xs, cs: O.log_prob(stack, state, xs, cs, RowNoise(lo, hi))
    sh = ShardedLogProb(local, mixtures=1)
    got = sh.log_prob(x, ctx)
    got2 = sh.log_prob_local(x[lo:hi], ctx[lo:hi], B)
    full = O.log_prob(stack, state, x, ctx, RowNoise(0, B))
    ok = torch.allclose(got, full, rtol=1e-6, atol=1e-5) and torch.equal(got, got2) and got.shape == (B, 1)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [10, 7])
def test_world2_gloo_shard_and_gather(B):
    port = 29500 + (os.getpid() % 2000) + B
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, B, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_bounds_cover_batch():
    from contextflow_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _grad_worker(rank, world, port, B, ret):
    """Training exchange: each rank back-propagates the mean loss of ITS slice (autograd over the oracle stands in for the per-rank
    CUDA backward), GradAllReduce combines; result must equal the single-process gradient of the global-batch mean loss."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from contextflow_b200 import synth
    from contextflow_b200.sharded import GradAllReduce, shard_bounds
    from oracle import flow_oracle as O
    from tests.golden.cases import CASES
    from tests.helpers import golden_state, load_golden
    case = dict(CASES['msl_conv_gen'], B=B)
    stack, state = golden_state(load_golden('msl_conv_gen'), case)
    names = [k for k, v in state.items() if v.is_floating_point() and not k.endswith('initialized')]
    params = [torch.nn.Parameter(state[k].clone()) for k in names]
    x, ctx = synth.make_inputs(case['conf'], B, 'gshard')
    eps = synth.normal('gshard:eps', (B, 1, 8, 1))

    class RowNoise:
        def __init__(self, lo, hi): self.lo, self.hi = lo, hi
        def randn(self, shape): return eps[self.lo:self.hi]
        def rand(self, shape): raise AssertionError

    def loss_grads(lo, hi):
        st = dict(state)
        for k, p in zip(names, params):
            p.grad = None; st[k] = p
        logp = O.log_prob(stack, st, x[lo:hi], ctx[lo:hi], RowNoise(lo, hi))
        cost, _, _ = O.training_loss(logp, None, case['conf']['data_size'], 1e2, criterion=False)
        cost.backward()
    lo, hi = shard_bounds(B, world, rank)
    loss_grads(lo, hi)
    unused = torch.nn.Parameter(torch.ones(3))                 # receives no gradient on any rank: must stay grad=None (AdamW skips it)
    GradAllReduce(params + [unused])(hi - lo, B)
    got = [None if p.grad is None else p.grad.clone() for p in params]
    loss_grads(0, B)
    ok = all((g is None) if p.grad is None else torch.allclose(g, p.grad, rtol=1e-4, atol=1e-5 * float(p.grad.abs().max()) + 1e-9)
             for g, p in zip(got, params))
    ret[rank] = bool(ok) and unused.grad is None and sum(p.grad is not None for p in params) >= 20
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [8, 7])
def test_world2_gloo_gradient_allreduce(B):
    port = 31500 + (os.getpid() % 2000) + B
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, port, B, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _actnorm_worker(rank, world, port, B, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from contextflow_b200 import synth
    from contextflow_b200.sharded import shard_bounds, sharded_actnorm_stats
    from oracle import flow_oracle as O
    x = synth.normal('anshard', (B, 6, 4, 3)) * 2.5 + synth.uniform('anshard:m', (1, 6, 1, 1)) * 4
    lo, hi = shard_bounds(B, world, rank)
    mean, logstd = sharded_actnorm_stats(x[lo:hi], local_stats=O.actnorm_init)
    ref_mean, ref_logstd = O.actnorm_init(x)
    ret[rank] = bool(torch.allclose(mean, ref_mean, rtol=1e-5, atol=1e-6) and torch.allclose(logstd, ref_logstd, rtol=1e-5, atol=1e-6))
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [9, 2])
def test_world2_gloo_actnorm_init_uses_global_batch_statistics(B):
    """the sharded first batch initialises ActNorm with the statistics of the global batch (actnorm.py:28-35)."""
    port = 33500 + (os.getpid() % 2000) + B
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_actnorm_worker, args=(2, port, B, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
