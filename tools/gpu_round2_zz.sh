set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_mirror.py tests/test_gpu_full_size.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02zz_pytest_sched.log 2>&1; tail -4 gpurun_out/r02zz_pytest_sched.log
for rep in 1 2; do
for cfg in "32 8" "26 1" "22 1"; do
  set -- $cfg
  timeout 300 python tools/search_shard_probe.py --L $1 --world $2 2>> gpurun_out/r02zz_probe.err >> gpurun_out/r02zz_shard_probe.jsonl
done
timeout 300 python tools/search_call_probe.py --L 20 2>> gpurun_out/r02zz_probe.err >> gpurun_out/r02zz_call_probe.jsonl
done
python - <<'P'
import json
for l in open("gpurun_out/r02zz_shard_probe.jsonl"):
    d = json.loads(l)
    print(d["L"], d["world"], {r: round(min(v["ms"][1:]), 4) for r, v in d["ranks"].items()})
for l in open("gpurun_out/r02zz_call_probe.jsonl"):
    d = json.loads(l)
    print(d["L"], d["bare_c_abi"], d["python_engine"])
P
tail -3 gpurun_out/r02zz_probe.err
