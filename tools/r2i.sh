#!/bin/bash
# conditioner epilogue: shared-space addressing + packed fp32 math: tests, per-level micro-bench (pipe on / off), bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_cond_tc.py -x -q -m gpu > gpurun_out/r2i_tc.log 2>&1; echo "tc tests rc=$?"; tail -c 300 gpurun_out/r2i_tc.log
for pipe in 0 1; do
  CFPP_TC_PIPE=$pipe timeout 300 python tools/bench_conv_cond.py 8192,16,16,16 8192,32,8,8 8192,64,4,4 > gpurun_out/r2i_cc_pipe$pipe.jsonl 2> gpurun_out/r2i_cc.err; echo "conv_cond bench pipe=$pipe rc=$?"
done
python - <<'P'
import json
for pipe in (0, 1):
    for l in open(f'gpurun_out/r2i_cc_pipe{pipe}.jsonl'):
        d = json.loads(l); c = d['cta0_cycles_per_tile']
        print('pipe', pipe, 'occ', d['plan']['occ'], d['shape'][1:], 'ms', d['tc_ms'], d['tc_TFLOPs'], 'issue', c['mma_issue'], 'wait_ops', c['mma_wait_ops'], 'epi', [c[k] for k in ('xform', 'epi1', 'epi2', 'epi3')], 'waits', [c[k] for k in ('wait_x0', 'wait_S1', 'wait_S2', 'wait_S3')], 'tot', d['cta0_total_per_tile'])
P
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:9]:
    print('   ', n, v['ms_per_step'], {k: s['ms_per_launch'] for k, s in v.get('by_shape', {}).items()})
P
