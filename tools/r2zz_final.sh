#!/bin/bash
# end-of-round validation (final session): smoke, full GPU suite, default bench line, reference arm, training step (default and CFPP_GMM_FAST=1)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2zz_smoke.log
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2zz_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2zz_tests.log
t0=$(date +%s); python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; echo "bench rc=$? elapsed $(( $(date +%s) - t0 )) s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2zz_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['parity_at_bench_batch'].get('ok'), d['roofline']['frac'], d['clocks'])
for k in ('secondary','tertiary'): print(k, round(d[k]['value']), d[k]['ms_per_step'], d[k]['parity_at_bench_batch'].get('ok'), d[k]['roofline'].get('frac'))
print(d['torch_cuda_baseline']['speedup_over_torch_cuda'], d['cpu_baseline']['value'])
P
t0=$(date +%s); python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zz_bench_reference.json 2> gpurun_out/r2zz_bench_reference.err; echo "reference arm rc=$? elapsed $(( $(date +%s) - t0 )) s"; tail -c 400 gpurun_out/r2zz_bench_reference.json
for f in 0 1; do
CFPP_GMM_FAST=$f timeout 600 python tools/bench_training.py --workload cfg2 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2zz_train_cfg2_fast$f.json 2> gpurun_out/r2zz_train.err; echo "train fast=$f rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2zz_train_cfg2_fast$f.json').read().strip().splitlines()[-1])
print('cfg2 training fast=$f', d['value'], d['ms_per_step'], d.get('loss'))
P
done
timeout 600 python tools/bench_training.py --workload cfg1 --batch 8192 --steps 5 --warmup 2 --graph > gpurun_out/r2zz_train_cfg1.json 2>> gpurun_out/r2zz_train.err; echo "train cfg1 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2zz_train_cfg1.json').read().strip().splitlines()[-1]); print('cfg1 training', d['value'], d['ms_per_step'])"
