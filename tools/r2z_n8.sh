#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2z_bench_n$N.json 2> gpurun_out/r2z_bench_n$N.err; echo "bench n$N rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2z_bench_n$N.json').read().strip().splitlines()[-1])
print('n$N', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['parity_at_bench_batch'].get('ok'))
for k in ('secondary','tertiary'): print(k, round(d[k]['value']), d[k]['ms_per_step'])
P
