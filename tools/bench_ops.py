"""Per-shape micro-benchmarks of the non-conditioner kernels at the cfg2 level shapes (run under gpurun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contextflow_b200 import ops

dev = 'cuda'
B = int(os.environ.get('B', 8192))
which = sys.argv[1:] or ['conv1x1', 'linear', 'gmm', 'coupling', 'actnorm', 'squeeze']


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


LEVELS = [(16, 16, 16), (32, 8, 8), (64, 4, 4)]
for D, H, W in LEVELS:
    HW = H * W
    xs = [torch.rand(B, D, H, W, device=dev) for _ in range(3)]
    it = [0]
    def nx():
        it[0] += 1; return xs[it[0] % 3]
    if 'conv1x1' in which:
        NN = torch.linalg.qr(torch.randn(D, D))[0].to(dev)
        lad = ops.slogdet(NN)
        cs = [torch.randn(B, D * D, device=dev) * 0.05 for _ in range(3)]
        lp = torch.randn(B, device=dev)
        t = timeit(lambda: ops.conv1x1(nx(), NN, lad, cs[it[0] % 3], lp, True))
        byt = 8.0 * B * D * HW + 4.0 * B * D * D
        t0 = timeit(lambda: ops.conv1x1(nx(), NN, lad))
        print(json.dumps({'op': 'conv1x1 ctx', 'D': D, 'HW': HW, 'us': round(t, 1), 'GBps': round(byt / t / 1e3, 1), 'shared_us': round(t0, 1),
                          'shared_GBps': round(8.0 * B * D * HW / t0 / 1e3, 1)}))
    if 'linear' in which:
        e = torch.randn(B, 20, device=dev)
        for name, K, N in [('CN conv1x1', 20, D * D), ('CN actnorm', 20, 2 * D), ('CN coup 1', 20, 2 * D), ('CN coup 2', 2 * D, 2 * D), ('CN coup 3', 2 * D, D)]:
            xin = torch.randn(B, K, device=dev)
            wt = torch.randn(K, N, device=dev); bb = torch.randn(N, device=dev)
            t = timeit(lambda: ops.linear(xin, wt, bb, relu=False))
            print(json.dumps({'op': 'linear ' + name, 'K': K, 'N': N, 'us': round(t, 1), 'out_GBps': round(4.0 * B * N / t / 1e3, 1)}))
    if 'coupling' in which:
        hs = [torch.randn(B, D, H, W, device=dev) for _ in range(3)]
        t = timeit(lambda: ops.coupling(nx(), hs[it[0] % 3]))
        print(json.dumps({'op': 'coupling', 'D': D, 'HW': HW, 'us': round(t, 1), 'GBps': round(12.0 * B * D * HW / t / 1e3, 1)}))
    if 'actnorm' in which:
        bt, bl = torch.randn(D, device=dev), torch.randn(D, device=dev) * 0.1
        cm = torch.randn(B, 2 * D, device=dev) * 0.1; lp = torch.randn(B, device=dev)
        t = timeit(lambda: ops.actnorm(nx(), bt, bl, cm, lp, float(HW), mode=1))
        print(json.dumps({'op': 'actnorm ctx', 'D': D, 'HW': HW, 'us': round(t, 1), 'GBps': round(8.0 * B * D * HW / t / 1e3, 1)}))
    if 'squeeze' in which:
        xq = torch.rand(B, D // 4, 2 * H, 2 * W, device=dev)
        t = timeit(lambda: ops.squeeze(xq, 2, 2))
        print(json.dumps({'op': 'squeeze', 'out': [D, H, W], 'us': round(t, 1), 'GBps': round(8.0 * B * D * HW / t / 1e3, 1)}))
    if 'gmm' in which and D in (16, 32, 64):
        Dg = D // 2 if D < 64 else D
        M, K = 10, 8
        mG = torch.randn(M, K, Dg, H, W, device=dev); sG = torch.randn(M, K, Dg, H, W, device=dev) * 0.1; wG = torch.randn(M, K, device=dev)
        xg = torch.randn(B, Dg, H, W, device=dev)
        ctx = torch.stack([torch.randint(0, 15, (B,)), torch.randint(0, 5, (B,))], 1).to(dev)
        width = 2 * M * K * Dg // 2
        tabs = [torch.randn(15, width, device=dev) * 0.05, torch.randn(5, width, device=dev) * 0.05]
        t = timeit(lambda: ops.gmm_logprob_ctxtab(xg, mG, sG, wG, ctx, [15, 5], tabs, 0.0))
        t0 = timeit(lambda: ops.gmm_logprob(xg, mG, sG, wG))
        fl = 3.0 * B * M * K * Dg * HW
        print(json.dumps({'op': 'gmm ctxtab', 'D': Dg, 'HW': HW, 'us': round(t, 1), 'TFLOPs': round(fl / t / 1e6, 2), 'noctx_us': round(t0, 1), 'noctx_TFLOPs': round(fl / t0 / 1e6, 2)}))
