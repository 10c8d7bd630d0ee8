#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_cond_tc.py tests/test_gpu_parity.py -q -m gpu --timeout 120 > gpurun_out/r2am_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2am_tests.log
for b in 256 1024 8192; do
timeout 300 python bench.py --batch $b --steps 30 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2am_bench_b$b.json 2> gpurun_out/r2am_bench.err; echo "b$b rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2am_bench_b$b.json').read().strip().splitlines()[-1])
print('b$b', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['e2e']['value'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:3]: print('   ', n, v['ms_per_step'], v['launches_per_step'], v.get('by_shape') and {k:x['ms_per_launch'] for k,x in v['by_shape'].items()})
P
done
