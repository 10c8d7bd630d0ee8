#!/bin/bash
mkdir -p gpurun_out
for v in "tc4:" "gen_fma:CFPP_VIT_GENERAL=1 CFPP_VIT_ATTN=fma" "gen_tc:CFPP_VIT_GENERAL=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --secondary= > gpurun_out/r2ah_cfg4_$name.json 2> gpurun_out/r2ah.err; echo "cfg4 $name rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/r2ah_cfg4_$name.json').read().strip().splitlines()[-1])
print('cfg4 $name', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['roofline'].get('frac'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:2]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
done
