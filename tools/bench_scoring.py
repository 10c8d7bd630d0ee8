"""Anomaly scoring end to end on one B200 (BASELINE configs[3]: SMAP-shaped 25-channel windows, ViT generalist): the raw (rows, 25) series
is uploaded ONCE from pinned host memory, every sliding window is generated on the device in the model layout (cfpp_windows_fwd), scored
(log_prob), reduced by the score epilogue (experiment_ad.py:262-278: dim_inv scale, NaN -> 0, theta(logp[:, -1])) and the scores are
read back.  One JSON line: windows/s with every host<->device byte inside the timed region.   python tools/bench_scoring.py [--rows N]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from contextflow_b200 import builder, ops, synth
from contextflow_b200.windows import WindowedScorer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=2_000_000); ap.add_argument('--batch', type=int, default=131072); ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    conf = synth.CONFIGS['cfg4']
    model = builder.build_named(conf)
    sd = model.state_dict(); synth.fill_state(sd, 'bench'); model.load_state_dict(sd)
    model = model.to(dev).eval()
    D, L = conf['data_size'][0], conf['data_size'][1]
    series = torch.rand(a.rows, D).pin_memory()
    dim_inv = 1.0 / float(D * L)
    out_host = torch.empty(a.rows, dtype=torch.float32).pin_memory()

    def run():
        sc = WindowedScorer(series, L, device=dev)                 # H2D of the raw series: rows * D * 4 bytes
        with torch.no_grad():
            for b0 in range(0, len(sc), a.batch):
                nb = min(a.batch, len(sc) - b0)
                x = sc.windows(b0, nb)
                ctx = torch.full((nb, 1), 7, device=dev, dtype=torch.int64)
                ep = ops.score_epilogue(model.log_prob(x, context=ctx), dim_inv, want=('last',))
                out_host[b0:b0 + nb].copy_(ep['last'], non_blocking=True)
        torch.cuda.synchronize()
    run()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        run()
    dt = (time.perf_counter() - t0) / a.reps
    print(json.dumps({'metric': 'anomaly_scoring_windows_per_sec_e2e', 'value': a.rows / dt, 'unit': 'windows/s', 'n_gpus': 1, 'rows': a.rows, 'window': L,
                      'channels': D, 'batch': a.batch, 'seconds': dt, 'h2d_bytes': a.rows * D * 4, 'd2h_bytes': a.rows * 4,
                      'h2d_bytes_if_windows_were_materialised_on_the_host': a.rows * D * L * 4,
                      'path': 'pinned series H2D once -> cfpp_windows_fwd -> log_prob (tcgen05 ViT) -> cfpp_score_epilogue -> D2H scores',
                      'finite': bool(torch.isfinite(out_host).all())}))


if __name__ == '__main__':
    main()
