#!/bin/bash
# round 2, GPU pass b: re-run the tests that failed in pass a, then ncu evidence: launch list + --set full captures
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_training.py -q -m gpu -k "bench_batch or graph_replay or every_context or experiment_loops or graphed or fallback" > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -c 1500 gpurun_out/r2b_tests.log
BENCH="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-parity --secondary ''"
$BENCH > gpurun_out/r2b_plain.log 2>&1 || { echo plain bench failed; tail -c 2000 gpurun_out/r2b_plain.log; exit 1; }
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv $BENCH > gpurun_out/r2b_ncu_launches.log 2>&1
for skip in 0 4 8; do
  CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_cond_tc_kernel --launch-skip $skip -c 1 \
    -f -o gpurun_out/r2b_conv_cond_tc_l$skip $BENCH > gpurun_out/r2b_ncu_cc$skip.log 2>&1
done
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv1x1 --launch-skip 0 -c 12 \
  -f -o gpurun_out/r2b_conv1x1 $BENCH > gpurun_out/r2b_ncu_c1.log 2>&1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"coupling_kernel|ctx_encode_batch|cn_batch|tile_kernel" -c 8 \
  -f -o gpurun_out/r2b_misc $BENCH > gpurun_out/r2b_ncu_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
