#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu --timeout 120 > gpurun_out/r2aw_train_tests.log 2>&1; echo "training tests rc=$?"; tail -4 gpurun_out/r2aw_train_tests.log
timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2aw_train_cfg2.json 2> gpurun_out/r2aw_train.err; echo "train rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2aw_train_cfg2.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d.get('loss'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:8]: print('   ', n, v['ms_per_step'], v['launches_per_step'], v.get('TFLOPps'))
P
