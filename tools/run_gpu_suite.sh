# full GPU suite + bench (run under gpurun)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG:-cur}.json 2> gpurun_out/bench_${TAG:-cur}.err; tail -c 3000 gpurun_out/bench_${TAG:-cur}.json; tail -5 gpurun_out/bench_${TAG:-cur}.err
