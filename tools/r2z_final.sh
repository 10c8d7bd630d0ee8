#!/bin/bash
# end-of-round validation: smoke, full GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zf_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2zf_smoke.log
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2zf_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2zf_tests.log
t0=$(date +%s); python bench.py > gpurun_out/r2zf_bench.json 2> gpurun_out/r2zf_bench.err; echo "bench rc=$? elapsed $(( $(date +%s) - t0 )) s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2zf_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['parity_at_bench_batch'].get('ok'), d['roofline']['frac'], d['clocks'])
for k in ('secondary','tertiary'): print(k, round(d[k]['value']), d[k]['ms_per_step'], d[k]['parity_at_bench_batch'].get('ok'), d[k]['roofline'].get('frac'))
print(d['torch_cuda_baseline']['speedup_over_torch_cuda'], d['cpu_baseline']['value'])
P
