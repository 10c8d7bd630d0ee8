#!/bin/bash
# round-2 profiling pass (B200_PROFILING.md): plain bench first, then the launch list of the same command, then `--set full` captures of the
# kernels the bench line's rooflines name.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2p_bench.json
python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2p_plain.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2p_ncu_launches.log 2>&1; echo "launch list rc=$?"
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"conv_cond_tc_kernel|conv1x1_ctx_kernel|coupling_kernel|tile_kernel" -c 16 \
  -o gpurun_out/r2p_cfg2_full python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2p_ncu_full.log 2>&1; echo "set full cfg2 rc=$?"
python bench.py --workload cfg4 --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2p_plain_cfg4.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"vit_tc4_kernel" -c 2 \
  -o gpurun_out/r2p_cfg4_full python bench.py --workload cfg4 --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2p_ncu_full_cfg4.log 2>&1; echo "set full cfg4 rc=$?"
ls -la gpurun_out/r2p_*
