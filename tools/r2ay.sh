#!/bin/bash
mkdir -p gpurun_out
for m in 31 24 8 16 25 26 28 27 29 30 15 23; do
CFPP_GMM_EXACT=$m timeout 300 python -m pytest tests/test_gpu_training.py -q -m gpu --timeout 120 -k "probsample or eyesample-False or cifar_vardeq-True or mixture" > gpurun_out/r2ay_tests_m$m.log 2>&1; echo "mode $m rc=$? $(tail -1 gpurun_out/r2ay_tests_m$m.log)"; grep -h "AssertionError:" gpurun_out/r2ay_tests_m$m.log | cut -c28-150
done
