# ncu evidence for the training and inverse directions (run under gpurun: TAG=r1x bash tools/ncu_training.sh)
set -x
mkdir -p gpurun_out
python tools/bench_training.py --batch 4096 --steps 1 --warmup 3 > gpurun_out/plain_train_${TAG:-cur}.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_train_${TAG:-cur}.csv python tools/bench_training.py --batch 4096 --steps 1 --warmup 3 > gpurun_out/ncu_train_${TAG:-cur}.log 2>&1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv2d_bwd_weight_kernel -c 4 \
  -o gpurun_out/conv2d_bwd_weight_${TAG:-cur} python tools/bench_training.py --batch 4096 --steps 1 --warmup 3 > gpurun_out/ncu_full_train_${TAG:-cur}.log 2>&1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:coupling_inv_kernel -c 4 \
  -o gpurun_out/coupling_inv_${TAG:-cur} python tools/bench_inverse.py --steps 1 > gpurun_out/ncu_full_inv_${TAG:-cur}.log 2>&1
ls -la gpurun_out/*.ncu-rep
