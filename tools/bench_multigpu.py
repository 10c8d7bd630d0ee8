"""Single-process N-GPU log_prob throughput (SURVEY §8e process model): weak (fixed rows per GPU) and strong (fixed global batch) scaling of
FlowSequential.log_prob through contextflow_b200/multigpu.py, host-timed around torch.cuda.synchronize() with CUDA events on the caller's
device.  python tools/bench_multigpu.py [--workload cfg2] [--per-gpu 8192] [--global-batch 65536]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contextflow_b200 import builder, synth

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='cfg2')
ap.add_argument('--per-gpu', type=int, default=8192)
ap.add_argument('--global-batch', type=int, default=65536)
ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()
os.environ['CFPP_CUDA_GRAPHS'] = '1'
conf = synth.CONFIGS[a.workload]
model = builder.build_named(conf)
sd = model.state_dict(); synth.fill_state(sd, 'bench'); model.load_state_dict(sd)
model = model.to('cuda:0').eval()
ndev = torch.cuda.device_count()


def run(B, devices):
    model.enable_multi_gpu(True, devices=[torch.device('cuda', i) for i in range(devices)], min_rows=256)
    xs = [tuple(t.to('cuda:0') for t in synth.make_inputs(conf, B, f'b{i}')) for i in range(2)]
    with torch.no_grad():
        for i in range(3):
            model.log_prob(*xs[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(a.steps):
            out = model.log_prob(*xs[i % 2])
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.steps
    ms = e0.elapsed_time(e1) / a.steps
    return dict(devices=devices, batch=B, ms_per_step=round(ms, 4), wall_ms=round(wall * 1e3, 4), samples_per_s=round(B / (ms * 1e-3)))


ns = [n for n in (1, 2, 4, 8) if n <= ndev]
rec = dict(workload=a.workload, form='single process, N GPUs (ReplicatedLogProb)', gpus_visible=ndev,
           weak=[run(a.per_gpu * n, n) for n in ns], strong=[run(a.global_batch, n) for n in ns])
print(json.dumps(rec))
