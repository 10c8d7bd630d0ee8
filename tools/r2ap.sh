#!/bin/bash
# register-tiled 3x3 backward-data route: op test first, then the whole GPU suite, then the cfg2 training step A/B (CFPP_BWD_DATA3=0 = generic tile)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py -q -m gpu -k "conv2d_family" --timeout 120 > gpurun_out/r2ap_conv.log 2>&1; echo "conv tests rc=$?"; tail -15 gpurun_out/r2ap_conv.log
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2ap_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ap_tests.log
for v in 1 0; do
CFPP_BWD_DATA3=$v timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2ap_train_cfg2_v$v.json 2> gpurun_out/r2ap_train_v$v.err; echo "train v=$v rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2ap_train_cfg2_v$v.json').read().strip().splitlines()[-1])
print('cfg2 bwd_data3=$v', d['value'], d['ms_per_step'], d.get('loss'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:8]: print('   ', n, v['ms_per_step'], v['launches_per_step'], v.get('TFLOPps'))
P
done
