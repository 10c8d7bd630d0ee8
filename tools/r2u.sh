#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_cond_tc.py -x -q -m gpu --timeout 60 > gpurun_out/r2u_tc.log 2>&1; echo "tc tests rc=$?"; tail -2 gpurun_out/r2u_tc.log
for mode in auto split; do
CFPP_CONV_COND=$mode timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2u_bench_$mode.json 2> gpurun_out/r2u_bench.err; echo "bench $mode rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2u_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:4]:
    print('   ', n, v['ms_per_step'], {k: s['ms_per_launch'] for k, s in v.get('by_shape', {}).items()})
P
done
