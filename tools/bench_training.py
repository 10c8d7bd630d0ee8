"""Measure one training step (experiment_ad.py:204-213: log_prob, loss, backward, torch AdamW step) of the context-free conv stack on
one B200 (or N with torchrun: per-rank slices + GradAllReduce), per-kernel CUDA-event table.  One JSON line on stdout.
python tools/bench_training.py [--workload cfg1] [--batch 4096] [--steps 10]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
import torch.distributed as dist
from contextflow_b200 import _cabi, builder, ops, synth
from contextflow_b200.sharded import GradAllReduce


def reference_arm(a):
    """The training step as the reference computes it: its eager torch op sequence (oracle/, pinned to the reference's gradients by
    tests/test_oracle_golden_training.py), torch autograd, torch AdamW -- on the GPU (torch-on-CUDA baseline) or the host cores."""
    import contextlib, time
    from oracle import flow_oracle as O
    conf = synth.variant(a.workload, enc_type=a.enc_type) if a.enc_type else synth.CONFIGS[a.workload]
    stack = O.build_stack(conf['cfg'], conf['data_size'], conf['mixtures'], conf['contexts'])
    model = builder.build_named(conf)
    state = model.state_dict(); synth.fill_state(state, 'bench')
    dev = a.ref_device
    O.DEVICE = dev
    state = {k: v.to(dev) for k, v in state.items()}
    names = [k for k, p in model.named_parameters() if p.requires_grad]
    params = []
    for k in names:
        state[k] = torch.nn.Parameter(state[k]); params.append(state[k])
    opt = torch.optim.AdamW(params, lr=1e-4)
    B, M = a.batch, conf['mixtures']
    C, H, W = conf['data_size']
    torch.manual_seed(99)
    x = (torch.randint(0, 256, (B, C, H, W)).float() if conf['image'] else torch.rand(B, C, H, W)).to(dev)
    ctx = torch.stack([torch.randint(0, k, (B,)) for k in conf['contexts']], 1).to(dev); gt = torch.randint(0, M, (B,)).to(dev)

    class TorchNoise:
        def rand(self, shape): return torch.rand(shape, device=dev)
        def randn(self, shape): return torch.randn(shape, device=dev)
    sync = torch.cuda.synchronize if dev == 'cuda' else (lambda: None)

    def step():
        opt.zero_grad(set_to_none=True)
        with (torch.device('cuda') if dev == 'cuda' else contextlib.nullcontext()):
            logp = O.log_prob(stack, state, x, ctx, TorchNoise())
            cost, _, _ = O.training_loss(logp, gt, conf['data_size'], 1e-2, True, None)
        cost.backward(); opt.step()
        return cost
    for _ in range(max(1, min(a.warmup, 2))):
        step()
    sync(); t0 = time.perf_counter()
    for _ in range(a.steps):
        cost = step()
    sync(); dt = time.perf_counter() - t0
    print(json.dumps({'metric': 'flow_training_step_samples_per_sec', 'impl': 'reference', 'value': B * a.steps / dt, 'unit': 'samples/s', 'n_gpus': 1,
                      'workload': a.workload + (f' (enc_type={a.enc_type})' if a.enc_type else ''), 'batch_per_gpu': B, 'ms_per_step': 1e3 * dt / a.steps, 'loss': float(cost.item()),
                      'how': f'oracle op sequence + torch autograd + AdamW, eager, device={dev}' + (f' ({torch.cuda.get_device_name(0)})' if dev == 'cuda' else f' ({torch.get_num_threads()} threads)')}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg1'); ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=10); ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--enc-type', default=None, help='override the workload encoder type, e.g. uniform (the reference default): '
                                                     'cfg2 with --enc-type uniform is the specialist whose training kernels exist')
    ap.add_argument('--graph', action='store_true', help='capture forward + loss + backward in one CUDA graph (GraphedTrainStep)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'],
                    help="reference: the reference's torch op sequence (oracle restatement) + torch autograd + AdamW, eager, on --ref-device")
    ap.add_argument('--ref-device', default='cuda', choices=['cpu', 'cuda'])
    a = ap.parse_args()
    if a.impl == 'reference':
        return reference_arm(a)
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local); dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    conf = synth.variant(a.workload, enc_type=a.enc_type) if a.enc_type else synth.CONFIGS[a.workload]
    model = builder.build_named(conf)
    sd = model.state_dict(); synth.fill_state(sd, 'bench'); model.load_state_dict(sd)
    model = model.to(dev).train()
    B, M = a.batch, conf['mixtures']
    C, H, W = conf['data_size']
    torch.manual_seed(99 + rank)
    batches = [((torch.randint(0, 256, (B, C, H, W)).float() if conf['image'] else torch.rand(B, C, H, W)).to(dev),
                torch.stack([torch.randint(0, k, (B,)) for k in conf['contexts']], 1).to(dev), torch.randint(0, M, (B,)).to(dev)) for _ in range(3)]
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4)
    sync = GradAllReduce(params)
    dim_inv = 1.0 / float(C * H * W)
    crit = nn.CrossEntropyLoss(); log_theta = nn.LogSigmoid()

    def loss_fn(m, x, c, gt):
        logp = dim_inv * m.log_prob(x, context=c)
        logp[logp != logp] = 0.0
        return crit(logp, gt) - 1e-2 * log_theta(torch.logsumexp(logp, -1)).mean()

    graphed = None
    if a.graph:
        from contextflow_b200.graphed import GraphedTrainStep
        graphed = GraphedTrainStep(model, loss_fn, *batches[0])

    def step(i):
        x, c, gt = batches[i % 3]
        if graphed is not None:
            cost = graphed(x, c, gt)
        else:
            opt.zero_grad(set_to_none=True)
            cost = loss_fn(model, x, c, gt)
            cost.backward()
        sync(B, B * world)
        opt.step()
        return cost

    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = os.environ.get('CFPP_PROFILE_RANGE') == '1'           # ncu --profile-from-start off: capture the timed steps only
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(a.steps):
        cost = step(i)
    e1.record(); torch.cuda.synchronize()
    if prof:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (_cabi.launch_count() - l0) / a.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    timer = ops.OpTimer(); ops.set_timer(timer)
    graphed_keep, graphed = graphed, None               # per-kernel events need eager launches
    for i in range(a.steps):
        step(i)
    torch.cuda.synchronize(); ops.set_timer(None)
    graphed = graphed_keep
    if rank == 0:
        summ = timer.summary()
        kern = {k: {'ms_per_step': round(v['ms'] / a.steps, 4), 'launches_per_step': v['n'] / a.steps,
                    'TFLOPps': round(v['flops'] / (v['ms'] / 1e3) / 1e12, 3) if v['flops'] else None,
                    'GBps': round(v['bytes'] / (v['ms'] / 1e3) / 1e9, 1) if v['bytes'] else None}
                for k, v in sorted(summ.items(), key=lambda kv: -kv[1]['ms'])}
        print(json.dumps({'metric': 'flow_training_step_samples_per_sec', 'value': world * B / (float(t.item()) / 1e3), 'unit': 'samples/s',
                          'n_gpus': world, 'workload': a.workload + (f' (enc_type={a.enc_type})' if a.enc_type else ''), 'batch_per_gpu': B, 'ms_per_step': float(t.item()), 'loss': float(cost.item()),
                          'libcfpp_launches_per_step': launches if not a.graph else 'captured in one CUDA graph (forward + loss + backward)', 'libcfpp_kernel_ms_per_step': round(sum(v['ms'] for v in summ.values()) / a.steps, 3),
                          'optimizer': 'torch.optim.AdamW (the reference builds it, model.py:289)', 'kernels': kern}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
