#!/bin/bash
# end-of-round dress rehearsal on 2 GPUs: full GPU suite, smoke, bench at N=1 / N=2 (torchrun) for both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2q_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2q_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2q_ref_n1.json 2> gpurun_out/r2q_ref_n1.err; echo "ref n1 rc=$?"; tail -c 300 gpurun_out/r2q_ref_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2q_bench_n2.json').read().strip().splitlines()[-1])
print('N=2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['n_gpus'], d.get('scaling'))
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2q_ref_n2.json 2> gpurun_out/r2q_ref_n2.err; echo "ref n2 rc=$?"; tail -c 200 gpurun_out/r2q_ref_n2.json
