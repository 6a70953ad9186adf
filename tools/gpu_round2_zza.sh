set -x
mkdir -p gpurun_out
for sh in 16 17 18 15; do
  export B200KF_TREE_SHIFT=$sh
  for L in 20 22 24; do
    timeout 300 python tools/search_call_probe.py --L $L --reps 500 2>> gpurun_out/r02zza_probe.err | sed "s/^{/{\"shift\": $sh, /" >> gpurun_out/r02zza_call_probe.jsonl
  done
  timeout 300 python tools/search_shard_probe.py --L 26 --world 1 2>> gpurun_out/r02zza_probe.err | sed "s/^{/{\"shift\": $sh, /" >> gpurun_out/r02zza_shard_probe.jsonl
done
python - <<'P'
import json
for l in open("gpurun_out/r02zza_call_probe.jsonl"):
    d = json.loads(l)
    print(d["shift"], d["L"], round(d["bare_c_abi"]["us_per_call"], 2))
for l in open("gpurun_out/r02zza_shard_probe.jsonl"):
    d = json.loads(l)
    print(d["shift"], d["L"], d["world"], {r: round(min(v["ms"][1:]), 4) for r, v in d["ranks"].items()})
P
tail -3 gpurun_out/r02zza_probe.err
