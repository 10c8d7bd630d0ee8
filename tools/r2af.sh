#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload cfg3 --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= --no-parity > gpurun_out/r2af_plain.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"vit_tc5_kernel" -c 12 \
  -o gpurun_out/r2af_cfg3_full python bench.py --workload cfg3 --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= --no-parity > gpurun_out/r2af_ncu_full.log 2>&1; echo "set full cfg3 rc=$?"
