#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2ao_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ao_tests.log
for wl in cfg1; do for tc in 0 1; do
CFPP_TRAIN_TC=$tc timeout 600 python tools/bench_training.py --workload $wl --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2ao_train_${wl}_tc$tc.json 2> gpurun_out/r2ao_train.err; echo "train $wl tc=$tc rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2ao_train_${wl}_tc$tc.json').read().strip().splitlines()[-1])
print('$wl tc=$tc', d['value'], d['ms_per_step'], d.get('loss'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:5]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
done; done
