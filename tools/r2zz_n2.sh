#!/bin/bash
# N = 2 sanity pass of the final build: multi-GPU tests, the bench line under torchrun, the data-parallel training step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multigpu.py -q -m gpu --timeout 300 > gpurun_out/r2zz_n2_tests.log 2>&1; echo "multigpu tests rc=$?"; tail -1 gpurun_out/r2zz_n2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --tertiary= > gpurun_out/r2zz_bench_n2.json 2> gpurun_out/r2zz_bench_n2.err; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2zz_bench_n2.json').read().strip().splitlines()[-1])
print(round(d['value']), d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['parity_at_bench_batch'].get('ok'), round(d['secondary']['value']))
P
