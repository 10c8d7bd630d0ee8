#!/bin/bash
mkdir -p gpurun_out
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv1x1_ctx_kernel" --launch-skip 8 -c 2 \
  -o gpurun_out/r2ak_c1_d64 python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= --no-parity > gpurun_out/r2ak_ncu.log 2>&1; echo "rc=$?"
