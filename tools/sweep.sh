# cfg5: large-batch throughput sweep at N=1 (run under gpurun: TAG=r1x bash tools/sweep.sh)
mkdir -p gpurun_out
: > gpurun_out/sweep_${TAG:-cur}.jsonl
for spec in "cfg2 256" "cfg2 4096" "cfg2 32768" "cfg4 256" "cfg4 8192" "cfg4 262144" "cfg1 256" "cfg1 32768" "cfg3 256" "cfg3 4096"; do
  set -- $spec
  timeout 600 python bench.py --workload $1 --batch $2 --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/sweep_${TAG:-cur}.jsonl 2>> gpurun_out/sweep_${TAG:-cur}.err || echo "{\"workload\": \"$1\", \"batch\": $2, \"failed\": true}" >> gpurun_out/sweep_${TAG:-cur}.jsonl
done
python - <<PY
import json
for l in open('gpurun_out/sweep_${TAG:-cur}.jsonl'):
    d = json.loads(l)
    if d.get('failed'): print(d); continue
    print(d['config']['workload'][:5], d['config']['batch_per_gpu'], round(d['value']), 'samples/s', round(d['ms_per_step'], 3), 'ms', 'e2e', round(d['e2e']['value']))
PY
