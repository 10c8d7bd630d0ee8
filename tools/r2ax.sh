#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 4 3 5 6 7; do
CFPP_GMM_EXACT=$m timeout 300 python -m pytest tests/test_gpu_training.py -q -m gpu --timeout 120 -k "golden or 3x_gate or fresh" > gpurun_out/r2ax_tests_m$m.log 2>&1; echo "mode $m rc=$? $(tail -1 gpurun_out/r2ax_tests_m$m.log)"; grep "^FAILED" gpurun_out/r2ax_tests_m$m.log | cut -c1-120
done
for m in 1 4 5; do
CFPP_GMM_EXACT=$m timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2ax_train_cfg2_m$m.json 2> gpurun_out/r2ax_train.err; echo "train rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2ax_train_cfg2_m$m.json').read().strip().splitlines()[-1])
print('cfg2 mode $m', d['value'], d['ms_per_step'], d.get('loss'), d['kernels']['gmm_ctx_train_bwd']['ms_per_step'], d['kernels']['gmm_ctx_train_fwd']['ms_per_step'])
P
done
