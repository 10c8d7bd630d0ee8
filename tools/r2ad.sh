#!/bin/bash
# A/B: 12 vs 16 epilogue warps in the one-CTA-per-SM conditioner kernels (libcfpp_e16.so), fused vs split route at B = 8192, cfg3 line
mkdir -p gpurun_out
E16=$PWD/contextflow_b200/libcfpp_e16.so
CFPP_LIB=$E16 timeout 300 python -m pytest tests/test_gpu_conv_cond_tc.py -x -q -m gpu --timeout 60 > gpurun_out/r2ad_tc_e16.log 2>&1; echo "tc tests (e16) rc=$?"; tail -2 gpurun_out/r2ad_tc_e16.log
for v in e12 e16; do
  if [ $v = e16 ]; then export CFPP_LIB=$E16; else unset CFPP_LIB; fi
  timeout 300 python tools/bench_conv_cond.py 8192,16,16,16 8192,32,8,8 8192,64,4,4 > gpurun_out/r2ad_cc_$v.jsonl 2> gpurun_out/r2ad_cc.err; echo "conv_cond bench $v rc=$?"
  python - <<P
import json
for l in open('gpurun_out/r2ad_cc_$v.jsonl'):
    d = json.loads(l); c = d['cta0_cycles_per_tile']
    print('$v pipe', d['plan']['pipe'], 'occ', d['plan']['occ'], d['shape'][1:], 'ms', d['tc_ms'], d['tc_TFLOPs'], 'issue', c['mma_issue'], 'wait_ops', c['mma_wait_ops'], 'epi', [c[k] for k in ('xform', 'epi1', 'epi2', 'epi3')], 'waits', [c[k] for k in ('wait_x0', 'wait_S1', 'wait_S2', 'wait_S3')], 'tot', d['cta0_total_per_tile'])
P
  for route in split fused; do
    CFPP_CONV_COND=$route timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2ad_bench_${v}_$route.json 2> gpurun_out/r2ad_bench.err; echo "bench $v $route rc=$?"
    python - <<P
import json
d=json.loads(open('gpurun_out/r2ad_bench_${v}_$route.json').read().strip().splitlines()[-1])
print('$v $route', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['gpu_launches'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:4]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
  done
done
unset CFPP_LIB
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --secondary= > gpurun_out/r2ad_bench_cfg3.json 2> gpurun_out/r2ad_bench_cfg3.err; echo "cfg3 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ad_bench_cfg3.json').read().strip().splitlines()[-1])
print('cfg3', round(d['value']), d['ms_per_step'], d['config'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:6]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
