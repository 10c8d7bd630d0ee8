#!/bin/bash
# round-2 profiling pass (B200_PROFILING.md): plain bench first, then the launch list of the same command, then `--set full` captures of the
# kernels the bench line's rooflines name.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2z_tests.log
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2z_bench.json
python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2z_plain.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2z_ncu_launches.log 2>&1; echo "launch list rc=$?"
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"conv_cond_tc_kernel|conv1x1_ctx_kernel|coupling_kernel|tile_kernel" -c 16 \
  -o gpurun_out/r2z_cfg2_full python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2z_ncu_full.log 2>&1; echo "set full cfg2 rc=$?"
ls -la gpurun_out/r2z_*
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_reference_arm.json 2> gpurun_out/r2z_reference_arm.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/r2z_reference_arm.json
timeout 300 python bench.py --batch 256 --steps 50 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2z_bench_b256.json 2> gpurun_out/r2z_bench_b256.err; echo "b256 rc=$?"
