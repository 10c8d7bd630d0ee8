#!/bin/bash
# fused conditioner + coupling as the default: full GPU suite, bench line, ncu --set full of the fused kernel (C = 16 level)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2t_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2t_tests.log
timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity_at_bench_batch'].get('ok'), d['gpu_launches'])
print('roofline', {k: d['roofline'][k] for k in ('kernel', 'frac', 'achieved', 'traffic')}); print('coupling', d['roofline_coupling'])
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:8]: print('   ', n, v['ms_per_step'])
print('secondary', round(d['secondary']['value']), d['secondary']['ms_per_step'])
P
python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2t_plain.log 2>&1 || exit 1
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2t_ncu_launches.log 2>&1; echo "launch list rc=$?"
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_cond_tc_kernel" -c 3 \
  -o gpurun_out/r2t_fused_full python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --secondary= > gpurun_out/r2t_ncu_full.log 2>&1; echo "set full rc=$?"
