#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s); python bench.py > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; echo "bench rc=$? elapsed $(( $(date +%s) - t0 )) s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ai_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['parity_at_bench_batch'].get('ok'))
for k in ('secondary','tertiary'): print(k, round(d[k]['value']), d[k]['ms_per_step'], d[k]['parity_at_bench_batch'].get('ok'), d[k]['roofline'].get('frac'), d[k]['config']['workload'])
P
