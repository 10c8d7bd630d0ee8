"""Micro-benchmark of the conv conditioner per BASELINE level (run under gpurun): tensor-core kernel vs FP32-FMA kernel,
CUDA-event timed over rotating inputs larger than L2, plus the in-kernel phase profile of CTA 0."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contextflow_b200 import ops, synth, _cabi

dev = 'cuda'
LEVELS = [(8192, 16, 16, 16), (8192, 32, 8, 8), (8192, 64, 4, 4), (8192, 8, 16, 16)]
if len(sys.argv) > 1:
    LEVELS = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for B, C, H, W in LEVELS:
    cin, ch, cout = C // 2, 2 * C, C
    g = torch.Generator().manual_seed(0)
    w1 = (torch.rand(ch, cin, generator=g) - 0.5).to(dev); w2 = ((torch.rand(ch, ch, 3, 3, generator=g) - 0.5) * 0.1).to(dev); w3 = (torch.rand(cout, ch, generator=g) - 0.5).to(dev)
    b1 = torch.zeros(ch, device=dev); b2 = torch.zeros(ch, device=dev); b3 = torch.zeros(cout, device=dev)
    xs = [torch.rand(B, C, H, W, device=dev) for _ in range(3)]
    pack = ops.conv_cond_tc_pack(w1, w2, w3, cin)
    pack = pack.repeat(int(os.environ.get('CFPP_TC_REPL', '1')))
    pk = (ops.pack_kmajor(w1), ops.pad_vec(b1), ops.pack_kmajor(w2.reshape(ch, -1)), ops.pad_vec(b2), ops.pack_kmajor(w3), ops.pad_vec(b3))
    flops = 2.0 * B * H * W * (cin * ch + ch * ch * 9 + ch * cout)

    def timeit(fn, n=12):
        for i in range(3): fn(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(xs[i % 3])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_tc = timeit(lambda x: ops.conv_cond_tc(x, cin, pack, b1, b2, b3, ch, H, W, 3, 3, cout))
    plan = ops.conv_cond_tc_last_plan()
    t_fma = timeit(lambda x: ops.conv_cond(x, cin, pk, H, W, 3, 3, cout))
    prof = torch.zeros(256, dtype=torch.int64, device=dev)
    _cabi.lib().cfpp_conv_cond_tc_set_profile(_cabi.vp(prof.data_ptr()))
    ops.conv_cond_tc(xs[0], cin, pack, b1, b2, b3, ch, H, W, 3, 3, cout)
    torch.cuda.synchronize()
    _cabi.lib().cfpp_conv_cond_tc_set_profile(None)
    pr = prof.cpu().tolist()
    slots = 148 * plan['occ']
    ntile0 = (plan['ntiles'] + slots - 1) // slots
    names = ['wait_x0', 'xform', 'wait_S1', 'epi1', 'wait_S2', 'epi2', 'wait_S3', 'epi3', 'mma_wait_w', 'mma_issue', 'mma_wait_ops', 'spare']
    print(json.dumps({'shape': [B, C, H, W], 'tc_ms': round(t_tc, 4), 'tc_TFLOPs': round(flops / t_tc / 1e9, 1), 'fma_ms': round(t_fma, 4),
                      'fma_TFLOPs': round(flops / t_fma / 1e9, 1), 'plan': plan,
                      'cta0_cycles_per_tile': {n: round(v / ntile0) for n, v in zip(names, pr)}, 'cta0_total_per_tile': round(sum(pr[:8]) / ntile0)}))
