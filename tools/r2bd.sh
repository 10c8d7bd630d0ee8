#!/bin/bash
# ncu --set full of the committed 3x3 backward-data kernel at the three cfg2 levels (B = 8192)
mkdir -p gpurun_out
ONLY=k3 VARIANTS=new REPS=1 WARM=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv2d_bwd_data3 -c 6 -o gpurun_out/r2bd_bwd_data3 python tools/bench_conv_train.py > gpurun_out/r2bd_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2bd_bwd_data3.ncu-rep
