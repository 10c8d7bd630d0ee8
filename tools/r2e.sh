#!/bin/bash
# round 2, GPU pass e: occupancy-2 fused conv1x1 kernel, one-exponential sigmoid flow: tests, bench, ncu --set full of the new kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv1x1" > gpurun_out/r2e_ops.log 2>&1; echo "ops rc=$?"; tail -c 600 gpurun_out/r2e_ops.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2e_parity.log 2>&1; echo "parity rc=$?"; tail -c 600 gpurun_out/r2e_parity.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary= > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python - <<'P'
import json
for f in ['r2e_bench']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d['parity_at_bench_batch'].get('ok'))
        for n,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:9]:
            print('   ', n, v['ms_per_step'], {k: s['ms_per_launch'] for k, s in v.get('by_shape', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
P
BENCH="python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline --no-parity --secondary="
CFPP_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv1x1_ctx_kernel|ctx_encode_flow_kernel|cn_batch_kernel" -c 14 \
  -f -o gpurun_out/r2e_c1ctx $BENCH > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2e_ncu.log
