#!/usr/bin/env python
"""Benchmark of the flow log-density hot path (BASELINE.json metric: flow log_prob samples/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm's CPU port (oracle/) on the host cores

A step = one log_prob pass over one synthetic batch of the workload (default: configs[1], the CIFAR-10C-shaped conv-coupling
specialist, onehot+vardeq, --contextflow).  Prints ONE JSON line on rank 0.
"""
import argparse, json, os, statistics, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'flow_log_prob_samples_per_sec'
UNIT = 'samples/s'
WORKLOADS = {'cfg1': 'mnist-r 1x32x32 conv generalist', 'cfg2': 'cifar10c 3x32x32 conv specialist onehot+vardeq contextflow',
             'cfg3': 'atm 38x144x1 trans specialist eye+argmax contextflow', 'cfg4': 'smap 25x8x1 trans generalist'}
DEFAULT_BATCH = {'cfg1': 8192, 'cfg2': 8192, 'cfg3': 1024, 'cfg4': 131072}
REF_BATCH = {'cfg1': 256, 'cfg2': 128, 'cfg3': 32, 'cfg4': 4096}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='samples per GPU per step (0 = workload default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ref-device', default='cpu', choices=['cpu', 'cuda'],
                    help='--impl reference only: cpu (the contract: host cores) or cuda (the same torch op sequence run eagerly on the GPU, '
                         'the "reference torch-on-CUDA" figure of the north star; informative)')
    ap.add_argument('--path', default='log_prob', choices=['log_prob', 'train', 'reverse'],
                    help='log_prob (default): the headline forward log-density path; train / reverse: the SURVEY §8(f) rows, measured by '
                         'tools/bench_training.py / tools/bench_inverse.py (their own JSON lines)')
    ap.add_argument('--eager', action='store_true', help='launch every kernel from Python instead of replaying the captured CUDA graph')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clocks + throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML through pynvml when importable (a
    sample per ~20 ms), else the recipe's nvidia-smi query line (one sample per ~0.3 s).  Samples carry a timestamp; summary() keeps
    those inside the timed window [t0, t1] and, when the window was too short for any, the nearest ones taken under the same load."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            power = 0.0
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, 'nvmlClocksEventReasonHwSlowdown', 0x8), getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), getattr(n, 'nvmlClocksEventReasonSwPowerCap', 0x4)]
        return [sm, self.max_sm, power] + [bool(mask & b) for b in bits]

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [v.strip() for v in out.split(',')]
        return [float(c[0]), float(c[1]), float(c[2])] + [v.lower().startswith('active') for v in c[3:7]]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                row = self._sample_nvml() if self.nvml is not None else self._sample_smi()
                self.rows.append((time.perf_counter(), row))
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.05)

    def summary(self, t0=None, t1=None):
        self.stop_flag.set()
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1)]
        window = 'timed region'
        if not rows and self.rows:                                      # window shorter than one sample: the samples closest to it (warm-up / e2e legs, same load)
            mid = 0.5 * (t0 + t1)
            rows = [r for _, r in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            window = 'nearest samples (timed region shorter than the sampling period)'
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[3 + i] for r in rows)]
        return {'sm_mhz': statistics.median(r[0] for r in rows), 'sm_max_mhz': rows[0][1], 'reasons': reasons, 'samples': len(rows),
                'power_w_max': max(r[2] for r in rows), 'window': window, 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def cpu_reference_run(workload, batch, steps, warmup, device='cpu'):
    """The reference algorithm (oracle/ restatement of the reference's torch op sequence, pinned to the reference by tests/golden) on all
    host cores -- or, device='cuda', the same eager op sequence on the GPU."""
    import contextlib
    import torch
    from contextflow_b200 import synth
    from oracle import flow_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    conf = synth.CONFIGS[workload]
    stack = O.build_stack(conf['cfg'], conf['data_size'], conf['mixtures'], conf['contexts'])
    from contextflow_b200 import builder
    model = builder.build_named(conf)
    state = model.state_dict(); synth.fill_state(state, 'bench')
    x, ctx = synth.make_inputs(conf, batch, 'bench')
    on_gpu = device == 'cuda'
    O.DEVICE = device
    if on_gpu:
        state = {k: v.cuda() for k, v in state.items()}
        x, ctx = x.cuda(), ctx.cuda()
    dev_ctx = torch.device('cuda') if on_gpu else contextlib.nullcontext()
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)

    ndev = 'cuda' if on_gpu else 'cpu'

    class TorchNoise:
        def rand(self, shape): return torch.rand(shape, device=ndev)
        def randn(self, shape): return torch.randn(shape, device=ndev)
    with torch.no_grad(), dev_ctx:
        for _ in range(warmup):
            O.log_prob(stack, state, x, ctx, TorchNoise())
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            O.log_prob(stack, state, x, ctx, TorchNoise())
        sync()
        dt = time.perf_counter() - t0
    how = f'torch CUDA eager fp32 on {torch.cuda.get_device_name(0)}' if on_gpu else f'torch CPU fp32, {torch.get_num_threads()} threads'
    return dict(value=batch * steps / dt, unit=UNIT, cores=cores, kind='port',
                sample=f'{steps} x log_prob of a {batch}-sample {workload} batch, {how}'), dt


def main():
    a = parse()
    if a.path != 'log_prob':                       # the "next" rows have their own measurement tools; same launch conventions (torchrun for N > 1)
        import runpy
        tool = 'bench_training.py' if a.path == 'train' else 'bench_inverse.py'
        argv = [tool, '--steps', str(a.steps)]
        if a.path == 'train':
            argv += ['--warmup', str(a.warmup), '--impl', a.impl, '--workload', a.workload]
            if a.batch:
                argv += ['--batch', str(a.batch)]
            if a.impl == 'reference':
                argv += ['--ref-device', a.ref_device]
        sys.argv = argv
        runpy.run_path(os.path.join(ROOT, 'tools', tool), run_name='__main__')
        return
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
    workload = a.workload

    if a.impl == 'reference':
        if rank != 0:
            return
        B = a.batch or REF_BATCH[workload]
        base, dt = cpu_reference_run(workload, B, a.steps, max(1, min(a.warmup, 2)), a.ref_device)
        print(json.dumps({'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'impl': 'reference', 'n_gpus': a.gpus, 'steps': a.steps,
                          'warmup': a.warmup, 'ms_per_step': 1e3 * dt / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': f'{workload}: {WORKLOADS[workload]}', 'batch_per_step': B, 'ref_device': a.ref_device},
                          'cpu_baseline': base, 'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    from contextflow_b200 import _cabi, builder, ops, synth
    from contextflow_b200.sharded import ShardedLogProb
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    saved_stdout = None
    if world > 1:
        # NCCL writes its version banner to stdout at communicator creation: keep this process's stdout for the ONE JSON line by
        # pointing fd 1 at stderr until the timed runs are over
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
    B = a.batch or DEFAULT_BATCH[workload]
    conf = synth.CONFIGS[workload]
    model = builder.build_named(conf)
    sd = model.state_dict(); synth.fill_state(sd, 'bench'); model.load_state_dict(sd)
    model = model.to(dev).eval()
    M = conf['mixtures']
    # synthetic inputs: NBUF rotating device-resident batches (HBM-resident inputs; the set is larger than L2) + pinned host copies
    NBUF = 3
    torch.manual_seed(1234 + rank)
    C, H, W = conf['data_size']
    def fresh():
        x = torch.randint(0, 256, (B, C, H, W)).float() if conf['image'] else torch.rand(B, C, H, W)
        ctx = torch.stack([torch.randint(0, k, (B,)) for k in conf['contexts']], 1)
        return x, ctx
    host = [tuple(t.pin_memory() for t in fresh()) for _ in range(NBUF)]
    devb = [(x.to(dev), c.to(dev)) for x, c in host]
    sharder = ShardedLogProb(lambda x, c: model.log_prob(x, c), M)
    from contextflow_b200.graphed import GraphedLogProb
    graphed = None if a.eager else GraphedLogProb(model)

    def run(x, c):
        """One log_prob pass: a replay of the captured CUDA graph (x / c are copied into its static input buffers -- device to
        device for resident inputs, pinned host to device for the end-to-end leg), or the eager launch sequence with --eager."""
        if graphed is not None:
            return graphed(x, c, clone=False)
        return model.log_prob(x.to(dev, non_blocking=True), c.to(dev, non_blocking=True))

    def step(i):
        x, c = devb[i % NBUF]
        with torch.no_grad():
            return sharder.gather(run(x, c), B * world)   # the path's only exchange: final score gather (no-op at N=1)

    def step_e2e(i, out_host):
        hx, hc = host[i % NBUF]
        with torch.no_grad():
            lp = sharder.gather(run(hx, hc), B * world)
            out_host.copy_(lp[:out_host.shape[0]], non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    for i in range(max(a.warmup, 3)):
        step(i)
    barrier()
    l0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    prof_range = os.environ.get('CFPP_PROFILE_RANGE') == '1'       # ncu --profile-from-start off: capture the timed steps only
    if prof_range:
        torch.cuda.profiler.start()
    t_begin = time.perf_counter()
    ev0.record()
    for i in range(a.steps):
        step(i)
    ev1.record()
    barrier()
    t_end = time.perf_counter()
    if prof_range:
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = _cabi.launch_count() - l0 if graphed is None else graphed.launches_per_replay(*devb[0]) * a.steps
    # per-kernel durations: the same steps launched eagerly with a CUDA-event pair around every libcfpp launch (events cannot be
    # read back from inside a replayed graph); same kernels, same inputs, right after the timed region
    timer = ops.OpTimer(); ops.set_timer(timer)
    with torch.no_grad():
        for i in range(a.steps):
            x, c = devb[i % NBUF]
            model.log_prob(x, c)
    barrier()
    ops.set_timer(None)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # end to end through the public API with host buffers (pinned H2D of x, ctx every step; D2H of every step's log-probs), timed on
    # the host clock.  Graph mode uses GraphedLogProb.stream: the copy of batch i+1 overlaps the replay of batch i.
    nout = B * world if rank == 0 else B
    outs_host = [torch.empty((nout, M), dtype=torch.float32).pin_memory() for _ in range(min(a.steps, 4))]
    gather = (lambda lp: sharder.gather(lp, B * world))

    def e2e_pass(n):
        if graphed is not None:
            with torch.no_grad():
                graphed.stream((host[i % NBUF] for i in range(n)), [outs_host[i % len(outs_host)] for i in range(n)], post=gather if world > 1 else None)
        else:
            for i in range(n):
                step_e2e(i, outs_host[i % len(outs_host)])
    e2e_pass(2)
    barrier()
    t0 = time.perf_counter()
    e2e_pass(a.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    clocks = sampler.summary(t_begin, t_end)
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
        tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        summ = timer.summary()
        total_ms = sum(v['ms'] for v in summ.values()) or 1.0
        kernels = {}
        for name, v in sorted(summ.items(), key=lambda kv: -kv[1]['ms']):
            sec = v['ms'] / 1e3
            kernels[name] = {'share': round(v['ms'] / total_ms, 4), 'ms_per_step': round(v['ms'] / a.steps, 4), 'launches_per_step': v['n'] / a.steps,
                             'GBps': round(v['bytes'] / sec / 1e9, 1) if sec > 0 else None, 'TFLOPps': round(v['flops'] / sec / 1e12, 3) if sec > 0 else None}
        top = next(iter(kernels))
        tv = summ[top]
        hbm_bound = top not in ('conv_cond_fwd', 'conv_cond_tc_fwd', 'vit_cond_fwd', 'gmm_logprob', 'gmm_logprob_ctxtab')
        if hbm_bound:
            ach = tv['bytes'] / (tv['ms'] / 1e3) / 1e9
            roof = {'kernel': top, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak, 'traffic': None, 'peak_source': hbm_src}
        else:
            ach = tv['flops'] / (tv['ms'] / 1e3) / 1e12
            traffic = None
            try:                                                   # DRAM bytes of one launch from the committed ncu --set full capture
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(top)
                if tj and tj.get('batch') == B and tj.get('workload') == workload:
                    traffic = tj['dram_bytes_per_launch']
            except Exception:
                pass
            roof = {'kernel': top, 'bound': 'tensor', 'achieved': ach, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': ach / tf_peak, 'traffic': traffic,
                    **({'issued_tensor_TFLOPps': 3.0 * ach, 'frac_issued': 3.0 * ach / tf_peak} if top == 'conv_cond_tc_fwd' else {}),
                    'peak_source': 'measured bf16 sustained (MEASURED_PEAKS.json); ' +
                                   ('achieved counts ALGORITHMIC flops: the fp32-faithful fp16-pair arithmetic issues 3 tensor products per algorithmic one, so its ceiling is 1/3 of this peak'
                                    if top == 'conv_cond_tc_fwd' else 'this kernel is an FP32-FMA path')}
        cv = summ.get('coupling_fwd')
        if cv:
            ach = cv['bytes'] / (cv['ms'] / 1e3) / 1e9
            roof_c = {'kernel': 'coupling_fwd', 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak, 'traffic': None}
        else:
            roof_c = None
        x0, c0 = host[0]
        line = {'metric': METRIC, 'value': world * B * a.steps / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
                'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': f'{workload}: {WORKLOADS[workload]}', 'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': f'dp{world} (batch-sharded, final all-gather of log-probs)',
                           'l2_policy': f'{NBUF} rotating input batches; per-layer working set {12 * B * C * H * W / 1e6:.0f} MB > 126 MB L2', 'weights': 'synthetic fill (synth.fill_state)',
                           'launch': 'eager (one Python call per kernel)' if graphed is None else 'CUDA graph replay of the whole log_prob (contextflow_b200/graphed.py)',
                           'kernel_timing': 'per-launch CUDA events in an eager pass over the same steps right after the timed region',
                           'e2e_mode': 'per-step blocking copies' if graphed is None else 'GraphedLogProb.stream: pinned H2D of batch i+1 on a copy stream overlaps the replay of batch i; D2H of every result'},
                'clocks': clocks, 'gpu_launches': launches,
                'e2e': {'value': world * B * a.steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': x0.numel() * 4 + c0.numel() * 8, 'd2h_bytes_per_step': B * world * M * 4},
                'roofline': roof, 'roofline_coupling': roof_c, 'kernels': kernels}
        if world == 1 and not a.no_cpu_baseline:
            # bounded sample of the same workload on the host cores: ~10 s of CPU work (probe one step, then as many as fit)
            probe, dt1 = cpu_reference_run(workload, REF_BATCH[workload], 1, 1)
            nsteps = max(2, min(60, int(10.0 / max(dt1, 1e-3))))
            base, _ = cpu_reference_run(workload, REF_BATCH[workload], nsteps, 0)
            line['cpu_baseline'] = base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
