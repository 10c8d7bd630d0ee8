"""Micro-benchmark of cfpp_cn_batch per job class at the cfg2 shapes (run under gpurun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contextflow_b200 import ops

dev = 'cuda'
B, K = int(os.environ.get('B', 8192)), 20


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def lin(k, n):
    return ops.pack_kmajor(torch.randn(n, k, device=dev) / k ** 0.5, 1), torch.randn(n, device=dev)


def make(kind, C):
    if kind == 'conv':
        return ops.cn_job([lin(K, C * C)], C)
    if kind == 'an':
        return ops.cn_job([lin(K, 2 * C)])
    return ops.cn_job([lin(K, 2 * C), lin(2 * C, 2 * C), lin(2 * C, C)])


for kind in ('conv', 'an', 'mlp', 'all'):
    for C in (16, 32, 64, 0):
        if (C == 0) != (kind == 'all'):
            continue
        jobs, keep = [], []
        kinds = ['conv', 'an', 'mlp'] if kind == 'all' else [kind]
        for kk in kinds:
            for CC in ((16, 32, 64) if C == 0 else (C,)):
                for _ in range(4):
                    j, k = make(kk, CC); jobs.append(j); keep.append(k)
        ins = [torch.randn(B, K, device=dev) for _ in jobs]
        t = timeit(lambda: ops.cn_batch(jobs, ins))
        print(json.dumps({'kind': kind, 'C': C, 'jobs': len(jobs), 'us': round(t, 1)}))
