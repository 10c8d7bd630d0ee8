#!/bin/bash
# ViT conditioner with four threads per token row: parity tests, cfg4 bench A/B against the one-thread-per-row kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -c 400 gpurun_out/r2m_tests.log
for v1 in 0; do
CFPP_VIT_TC_V1=$v1 timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --secondary= > gpurun_out/r2m_bench_cfg4_v1$v1.json 2> gpurun_out/r2m_bench_cfg4_v1$v1.err; echo "bench v1=$v1 rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2m_bench_cfg4_v1$v1.json').read().strip().splitlines()[-1])
print('v1=$v1', round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'), d['roofline'].get('frac'), d['roofline'].get('achieved'))
for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:5]:
    print('   ', n, v['ms_per_step'])
P
done
