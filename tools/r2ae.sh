#!/bin/bash
# four-threads-per-row general ViT kernel (vit_tc5): op-level parity, full-model parity, cfg3 A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py -q -m gpu --timeout 120 > gpurun_out/r2ae_vit.log 2>&1; echo "vit tests rc=$?"; tail -15 gpurun_out/r2ae_vit.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -k "cfg3 or atm or smd or cfg4" > gpurun_out/r2ae_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r2ae_parity.log
for v1 in 0; do
for attn in fma tc; do CFPP_VIT_ATTN=$attn timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --secondary= > gpurun_out/r2ae_bench_cfg3_v1$v1.json 2> gpurun_out/r2ae_bench_cfg3.err; echo "cfg3 v1=$v1 rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2ae_bench_cfg3_v1$v1.json').read().strip().splitlines()[-1])
print('cfg3 attn=$attn', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['roofline'].get('frac'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:3]: print('   ', n, v['ms_per_step'], v['launches_per_step'])
P
done
done
