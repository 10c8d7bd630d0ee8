# inverse-direction GPU tests + microbench (run under gpurun)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inverse.py -x -q 2>&1 | tail -25
timeout 600 python tools/bench_inverse.py --steps 10 > gpurun_out/bench_inverse_${TAG:-cur}.json 2> gpurun_out/bench_inverse_${TAG:-cur}.err; cat gpurun_out/bench_inverse_${TAG:-cur}.json; tail -5 gpurun_out/bench_inverse_${TAG:-cur}.err
