#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_boundary.py -q -m gpu --timeout 120 > gpurun_out/r2bc_train_tests.log 2>&1; echo "training+boundary tests rc=$?"; tail -2 gpurun_out/r2bc_train_tests.log
timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2bc_train_cfg2.json 2> gpurun_out/r2bc_train.err; echo "train rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2bc_train_cfg2.json').read().strip().splitlines()[-1])
ks=d['kernels']; print('cfg2 training', d['value'], d['ms_per_step'], d.get('loss'), {k: round(ks[k]['ms_per_step'],1) for k in ('conv2d_bwd_data','gmm_ctx_train_bwd','gmm_ctx_train_fwd','embed_scatter')})
P
