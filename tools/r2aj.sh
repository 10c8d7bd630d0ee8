#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -k "gmm or mixture or prior" > gpurun_out/r2aj_gmm.log 2>&1; echo "gmm tests rc=$?"; tail -2 gpurun_out/r2aj_gmm.log
timeout 300 python tools/bench_gmm.py > gpurun_out/r2aj_gmm.jsonl 2> gpurun_out/r2aj_gmm.err; echo "bench_gmm rc=$?"; cat gpurun_out/r2aj_gmm.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['gpu_launches'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:6]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
