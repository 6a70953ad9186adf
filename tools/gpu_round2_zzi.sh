set -x
mkdir -p gpurun_out
lscpu | grep -E "Model name|L3|L2|Socket|^CPU\(s\)" > gpurun_out/r02zzi_lscpu.txt; cat gpurun_out/r02zzi_lscpu.txt
for mb in 128 16 32 64 8; do
  B200KF_HOST_CHUNK_MB=$mb timeout 600 python bench.py --no-configs --no-cpu-baseline --no-search --no-traj --no-shared --steps 3 --warmup 3 2>> gpurun_out/r02zzi_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print(json.dumps({'chunk_mb': $mb, 'pinned_value': e['value'], 'pinned_ms': e['ms_per_call'], 'pageable': e.get('pageable_numpy')}))" >> gpurun_out/r02zzi_chunk.jsonl
done
cat gpurun_out/r02zzi_chunk.jsonl; tail -2 gpurun_out/r02zzi_bench.err
