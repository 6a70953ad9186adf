set -x
mkdir -p gpurun_out
nproc; grep -c processor /proc/cpuinfo
timeout 600 python -m pytest tests/test_gpu_run_batch.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "host or pageable" > gpurun_out/r02zw_pytest.log 2>&1; tail -4 gpurun_out/r02zw_pytest.log
for th in 0 4 8 16; do
  B200KF_STAGE_THREADS=$th timeout 600 python bench.py --no-configs --no-cpu-baseline --no-search --no-traj --no-shared --steps 3 --warmup 3 2>> gpurun_out/r02zw_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print(json.dumps({'stage_threads': $th, 'pinned_value': e['value'], 'pinned_ms': e['ms_per_call'], 'pageable': e.get('pageable_numpy')}))" >> gpurun_out/r02zw_pageable.jsonl
done
cat gpurun_out/r02zw_pageable.jsonl; tail -3 gpurun_out/r02zw_bench.err
