#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q -m gpu --timeout 300 > gpurun_out/r2z_n2_tests.log 2>&1; echo "multigpu tests rc=$?"; tail -3 gpurun_out/r2z_n2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/r2z_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2z_ref_n2.json 2> gpurun_out/r2z_ref_n2.err; echo "ref n2 rc=$?"; tail -c 200 gpurun_out/r2z_ref_n2.json
