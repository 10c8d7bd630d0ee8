#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py -q -m gpu -k "conv2d_family" --timeout 120 > gpurun_out/r2ar_conv.log 2>&1; echo "conv tests rc=$?"; tail -4 gpurun_out/r2ar_conv.log
ONLY=k3 timeout 300 python tools/bench_conv_train.py > gpurun_out/r2ar_conv_shapes.jsonl 2> gpurun_out/r2ar_conv_shapes.err; echo "shapes rc=$?"; cat gpurun_out/r2ar_conv_shapes.jsonl
ONLY=k3 VARIANTS=new REPS=1 WARM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_bwd_data3 -c 6 -o gpurun_out/r2ar_bwd_data3 python tools/bench_conv_train.py > gpurun_out/r2ar_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
