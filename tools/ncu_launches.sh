# launch list of the timed steps (eager launches), per B200_PROFILING.md; run under gpurun: TAG=r1x bash tools/ncu_launches.sh
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/plain_${TAG:-cur}.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_${TAG:-cur}.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/ncu_${TAG:-cur}.log 2>&1
tail -c 300 gpurun_out/ncu_${TAG:-cur}.log
