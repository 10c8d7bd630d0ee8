#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2ac_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['gpu_launches'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:10]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
