"""Measure the inverse (sampling) direction and the score epilogue on one B200: samples/s of FlowSequential.reverse / .sample for the
generalist stacks (the models whose inverse the reference can execute), per-kernel CUDA-event times with the HBM roofline of
coupling_inv, and the score epilogue's GB/s.  One JSON line per measurement on stdout.   python tools/bench_inverse.py [--steps 20]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from contextflow_b200 import builder, ops, synth


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6650.0)
    except Exception:
        return 6650.0


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def kernel_table(fn, steps, hbm):
    t = ops.OpTimer(); ops.set_timer(t)
    for _ in range(steps):
        fn()
    torch.cuda.synchronize(); ops.set_timer(None)
    out = {}
    for name, v in sorted(t.summary().items(), key=lambda kv: -kv[1]['ms']):
        sec = v['ms'] / 1e3
        out[name] = {'ms_per_step': round(v['ms'] / steps, 4), 'launches_per_step': v['n'] / steps,
                     'GBps': round(v['bytes'] / sec / 1e9, 1) if v['bytes'] and sec > 0 else None,
                     'hbm_frac': round(v['bytes'] / sec / 1e9 / hbm, 3) if v['bytes'] and sec > 0 else None}
    return out


def main():
    ap = argparse.ArgumentParser(); ap.add_argument('--steps', type=int, default=20)
    a = ap.parse_args()
    hbm = peaks()
    dev = torch.device('cuda', 0)
    for name, B in (('cfg1', 16384), ('cfg4', 131072)):
        conf = synth.CONFIGS[name]
        model = builder.build_named(conf)
        sd = model.state_dict(); synth.fill_state(sd, 'bench'); model.load_state_dict(sd)
        model = model.to(dev).eval()
        C, H, W = conf['data_size']
        x = (torch.randint(0, 256, (B, C, H, W)).float() if conf['image'] else torch.rand(B, C, H, W)).to(dev)
        ctx = torch.stack([torch.randint(0, k, (B,)) for k in conf['contexts']], 1).to(dev)
        with torch.no_grad():
            zs = [model(x, ctx)[0] for _ in range(3)]                  # 3 rotating latents: working set > L2
            i = [0]
            def rev():
                i[0] += 1
                return model.reverse(zs[i[0] % 3], ctx)
            if os.environ.get('CFPP_PROFILE_RANGE') == '1':
                torch.cuda.profiler.start(); rev(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
            ms_rev = timed(rev, a.steps)
            ms_fwd = timed(lambda: model.log_prob(x, ctx), a.steps)
            line = {'metric': 'flow_reverse_samples_per_sec', 'value': B / (ms_rev / 1e3), 'unit': 'samples/s', 'workload': name, 'batch': B,
                    'ms_per_step': ms_rev, 'forward_log_prob_ms': ms_fwd, 'kernels': kernel_table(rev, a.steps, hbm), 'hbm_peak_GBps': hbm}
            if model.dist.M >= 2:
                ms_s = timed(lambda: model.sample(B), a.steps)
                line['sample_ms_per_step'] = ms_s; line['sample_samples_per_sec'] = B / (ms_s / 1e3)
            print(json.dumps(line), flush=True)
        del model, zs
    B, M = 1 << 20, 10
    logp = torch.randn(B, M, device=dev) * 300 - 9000
    gt = torch.randint(0, M, (B,), device=dev)
    ms = timed(lambda: ops.score_epilogue(logp, 1.0 / 3072, gt, None), a.steps)
    byt = 4.0 * B * M * 2 + B * (4 * 3 + 8 + 8)
    print(json.dumps({'metric': 'score_epilogue', 'batch': B, 'M': M, 'ms': ms, 'GBps': byt / ms / 1e6, 'hbm_frac': byt / ms / 1e6 / hbm,
                      'note': 'includes the torch allocations of its outputs'}), flush=True)


if __name__ == '__main__':
    main()
