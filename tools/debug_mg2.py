import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CFPP_CUDA_GRAPHS'] = '1'
import torch
from contextflow_b200 import builder, synth
from contextflow_b200.multigpu import _peer_copy
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1); return round(1e3 * (time.perf_counter() - t0) / n, 3)
conf = synth.variant('cfg4', data_size=(24, 8, 1)); B = 32768
model = builder.build_named(conf)
sd = model.state_dict(); synth.fill_state(sd, 'mg'); model.load_state_dict(sd)
model = model.to('cuda:0').eval()
x, ctx = synth.make_inputs(conf, B, 'mg'); x, ctx = x.cuda(), ctx.cuda()
half = B // 2
with torch.no_grad():
    model.enable_multi_gpu(min_rows=256); model.log_prob(x, ctx)
    rep = model.__dict__['_replicated']._replicas[1]
    xh, ch = x[:half].contiguous(), ctx[:half].contiguous()
    x1, c1 = x[half:].to('cuda:1'), ctx[half:].to('cuda:1')
    dev0, dev1 = torch.device('cuda:0'), torch.device('cuda:1')
    def pipe(wait_ready, pull, push, wait_done, cat, own_stream=False):
        s1x = torch.cuda.Stream(dev1) if own_stream else None
        def f():
            cur = torch.cuda.current_stream(dev0)
            ready = torch.cuda.Event(); ready.record(cur)
            back = torch.empty((B - half, model.mixtures), device=dev0)
            with torch.cuda.device(dev1):
                s = torch.cuda.current_stream(dev1)
                if wait_ready: s.wait_event(ready)
                if pull:
                    xs = torch.empty_like(x1); cs = torch.empty_like(c1)
                    _peer_copy(xs, x[half:], s); _peer_copy(cs, ctx[half:], s)
                else:
                    xs, cs = x1, c1
                lp = rep._log_prob_single(xs, cs)
                if push: _peer_copy(back, lp, s)
                done = torch.cuda.Event(); done.record(s)
            a = model._log_prob_single(xh, ch)
            if wait_done: cur.wait_event(done)
            if cat: return torch.cat([a, back], 0)
        return f
    for cfg in [(0,0,0,0,0), (1,0,0,0,0), (1,0,0,1,0), (1,1,0,1,0), (1,0,1,1,0), (1,1,1,1,0), (1,1,1,1,1), (0,1,1,0,0), (0,1,0,0,0), (0,0,1,0,0)]:
        print(dict(zip(('wait_ready', 'pull', 'push', 'wait_done', 'cat'), cfg)), timeit(pipe(*cfg)), 'ms')
