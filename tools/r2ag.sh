#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2ag_tests.log
timeout 300 python bench.py --workload cfg3 --batch 8192 --steps 5 --warmup 3 --no-cpu-baseline --secondary= > gpurun_out/r2ag_bench_cfg3_b8192.json 2> gpurun_out/r2ag_bench_cfg3.err; echo "cfg3 b8192 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ag_bench_cfg3_b8192.json').read().strip().splitlines()[-1])
print('cfg3 b8192', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['roofline'].get('frac'))
P
python bench.py --workload cfg3 --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= --no-parity > gpurun_out/r2ag_plain.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"vit_tc5_kernel" -c 12 \
  -o gpurun_out/r2ag_cfg3_full python bench.py --workload cfg3 --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= --no-parity > gpurun_out/r2ag_ncu_full.log 2>&1; echo "set full cfg3 rc=$?"
