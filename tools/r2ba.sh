#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu --timeout 120 > gpurun_out/r2ba_train_tests.log 2>&1; echo "training tests rc=$?"; tail -2 gpurun_out/r2ba_train_tests.log
for f in 0 1; do
CFPP_GMM_FAST=$f timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2ba_train_cfg2_fast$f.json 2> gpurun_out/r2ba_train.err; echo "train fast=$f rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2ba_train_cfg2_fast$f.json').read().strip().splitlines()[-1])
ks=d['kernels']; print('cfg2 training fast=$f', d['value'], d['ms_per_step'], d.get('loss'), {k: round(ks[k]['ms_per_step'],1) for k in ('conv2d_bwd_data','gmm_ctx_train_bwd','gmm_ctx_train_fwd')})
P
done
