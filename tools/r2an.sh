#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_cond_tc.py -q -m gpu --timeout 120 > gpurun_out/r2an_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r2an_tc.log
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu --timeout 120 > gpurun_out/r2an_train.log 2>&1; echo "training tests rc=$?"; tail -3 gpurun_out/r2an_train.log
for tc in 0 1; do
CFPP_TRAIN_TC=$tc timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2an_train_cfg2_tc$tc.json 2> gpurun_out/r2an_train_cfg2.err; echo "train cfg2 tc=$tc rc=$?"; tail -c 500 gpurun_out/r2an_train_cfg2_tc$tc.json | head -c 500; echo
done
