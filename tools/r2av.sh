#!/bin/bash
mkdir -p gpurun_out
CFPP_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gmm_ctx -c 6 \
  -o gpurun_out/r2av_gmm_ctx python tools/bench_training.py --workload cfg2 --batch 8192 --steps 1 --warmup 2 > gpurun_out/r2av_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2av_ncu.log
ls -la gpurun_out/*.ncu-rep
