#!/bin/bash
# round 2, first GPU pass: new parity / graphed-training / boundary tests, then the whole suite, then the bench (both arms)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -x -q -m gpu -k "bench_batch or graph_replay or every_context or experiment_loops" > gpurun_out/r2a_new_tests.log 2>&1
echo "new parity tests rc=$?" | tee -a gpurun_out/r2a_new_tests.log
python -m pytest tests/test_gpu_training.py -x -q -m gpu -k "graphed or fallback" > gpurun_out/r2a_train_tests.log 2>&1
echo "graphed training tests rc=$?" | tee -a gpurun_out/r2a_train_tests.log
python -m pytest tests -q -m gpu > gpurun_out/r2a_all_tests.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r2a_all_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
echo "bench ref rc=$?"
python bench.py --steps 20 --warmup 5 --batch 256 --secondary '' --no-cpu-baseline > gpurun_out/r2a_bench_b256.json 2> gpurun_out/r2a_bench_b256.err
echo "bench b256 rc=$?"
tail -c 600 gpurun_out/r2a_new_tests.log; tail -c 600 gpurun_out/r2a_train_tests.log; tail -c 800 gpurun_out/r2a_all_tests.log
