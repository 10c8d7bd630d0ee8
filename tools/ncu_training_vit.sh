# ncu evidence for the ViT training step (run under gpurun: TAG=r1x bash tools/ncu_training_vit.sh)
set -x
mkdir -p gpurun_out
python tools/bench_training.py --workload cfg4 --batch 8192 --steps 1 --warmup 3 > gpurun_out/plain_train_vit_${TAG:-cur}.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_train_vit_${TAG:-cur}.csv python tools/bench_training.py --workload cfg4 --batch 8192 --steps 1 --warmup 3 > gpurun_out/ncu_train_vit_${TAG:-cur}.log 2>&1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rows_wgrad_kernel -c 3 \
  -o gpurun_out/rows_wgrad_${TAG:-cur} python tools/bench_training.py --workload cfg4 --batch 8192 --steps 1 --warmup 3 > gpurun_out/ncu_full_train_vit_${TAG:-cur}.log 2>&1
ls -la gpurun_out/*${TAG:-cur}*
