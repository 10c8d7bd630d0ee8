#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_cond_tc.py -x -q -m gpu --timeout 60 > gpurun_out/r2y_tc.log 2>&1; echo "tc tests rc=$?"; tail -2 gpurun_out/r2y_tc.log
for pipe in 0 1; do
CFPP_TC_PIPE=$pipe timeout 300 python tools/bench_conv_cond.py 8192,16,16,16 8192,32,8,8 8192,64,4,4 > gpurun_out/r2y_cc$pipe.jsonl 2> gpurun_out/r2y_cc.err; echo "conv_cond bench rc=$?"
python - <<P
import json
for l in open('gpurun_out/r2y_cc$pipe.jsonl'):
    d = json.loads(l); c = d['cta0_cycles_per_tile']
    print('pipe', d['plan']['pipe'], 'occ', d['plan']['occ'], d['shape'][1:], 'ms', d['tc_ms'], d['tc_TFLOPs'], 'issue', c['mma_issue'], 'wait_ops', c['mma_wait_ops'], 'epi', [c[k] for k in ('xform', 'epi1', 'epi2', 'epi3')], 'waits', [c[k] for k in ('wait_x0', 'wait_S1', 'wait_S2', 'wait_S3')], 'tot', d['cta0_total_per_tile'])
P
done
