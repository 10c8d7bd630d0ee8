# one `ncu --set full` capture of the dominant kernel (tcgen05 conv conditioner, C=16 level = first launch of a step) inside bench.py
set -x
mkdir -p gpurun_out
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_cond_tc_kernel -c 3 \
  -o gpurun_out/conv_cond_tc_${TAG:-cur} python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/ncu_full_${TAG:-cur}.log 2>&1
ls -la gpurun_out/conv_cond_tc_${TAG:-cur}.ncu-rep
