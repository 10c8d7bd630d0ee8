#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_training.py -q -m gpu --timeout 120 > gpurun_out/r2al_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2al_tests.log
timeout 300 python bench.py --batch 256 --steps 50 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2al_bench_b256.json 2> gpurun_out/r2al_bench_b256.err; echo "b256 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2al_bench_b256.json').read().strip().splitlines()[-1])
print('b256', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['e2e']['value'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:5]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
