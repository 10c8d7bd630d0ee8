#!/bin/bash
# round 2, GPU pass f: software-pipelined conditioner tiles (CFPP_TC_PIPE), multi-stage fused conv1x1: tests, per-level micro-bench A/B, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_cond_tc.py -x -q -m gpu > gpurun_out/r2f_tc.log 2>&1; echo "tc tests rc=$?"; tail -c 800 gpurun_out/r2f_tc.log
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2f_parity.log 2>&1; echo "ops+parity rc=$?"; tail -c 600 gpurun_out/r2f_parity.log
for pipe in 1 0; do
  CFPP_TC_PIPE=$pipe timeout 300 python tools/bench_conv_cond.py 8192,16,16,16 8192,32,8,8 8192,64,4,4 > gpurun_out/r2f_cc_pipe$pipe.jsonl 2> gpurun_out/r2f_cc_pipe$pipe.err; echo "conv_cond bench pipe=$pipe rc=$?"
done
python - <<'P'
import json
for pipe in (1, 0):
    for l in open(f'gpurun_out/r2f_cc_pipe{pipe}.jsonl'):
        d = json.loads(l); print('pipe', pipe, d['shape'], d['tc_ms'], d['tc_TFLOPs'], {k: d['plan'][k] for k in ('S', 'T1', 'T2', 'nstages', 'occ', 'pipe', 'smem_bytes')}, d['cta0_cycles_per_tile'], d['cta0_total_per_tile'])
P
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'P'
import json
for f in ['r2f_bench']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'))
        for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:9]:
            print('   ', n, v['ms_per_step'], {k: s['ms_per_launch'] for k, s in v.get('by_shape', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
P
