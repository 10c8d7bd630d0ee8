#!/bin/bash
# conditioner MMA-issue experiment: in-kernel phase counters with the epilogue work switched off (CFPP_TC_DBG)
mkdir -p gpurun_out
for cfg in "1 0" "0 2" "0 1"; do
  set -- $cfg; pipe=$1; occ=$2
  for dbg in 0 1 3; do
    CFPP_TC_PIPE=$pipe CFPP_TC_OCC=$occ CFPP_TC_DBG=$dbg timeout 300 python tools/bench_conv_cond.py 8192,16,16,16 8192,32,8,8 8192,64,4,4 > gpurun_out/r2g_p${pipe}_o${occ}_d${dbg}.jsonl 2> gpurun_out/r2g.err || tail -3 gpurun_out/r2g.err
  done
done
python - <<'P'
import json
for pipe, occ in ((1, 0), (0, 2), (0, 1)):
    for dbg in (0, 1, 3):
        for l in open(f'gpurun_out/r2g_p{pipe}_o{occ}_d{dbg}.jsonl'):
            d = json.loads(l); c = d['cta0_cycles_per_tile']
            print('pipe', pipe, 'occ', d['plan']['occ'], 'dbg', dbg, d['shape'][1:], 'ms', d['tc_ms'], 'S', d['plan']['S'], 'T1', d['plan']['T1'], 'T2', d['plan']['T2'], 'nst', d['plan']['nstages'],
                  'issue', c['mma_issue'], 'wait_ops', c['mma_wait_ops'], 'epi', [c[k] for k in ('xform', 'epi1', 'epi2', 'epi3')], 'waits', [c[k] for k in ('wait_x0', 'wait_S1', 'wait_S2', 'wait_S3')], 'tot', d['cta0_total_per_tile'])
P
