set -x
mkdir -p gpurun_out
for L in 20 16 24; do timeout 300 python tools/search_call_probe.py --L $L 2>> gpurun_out/r02zr_probe.err >> gpurun_out/r02zr_search_call_probe.jsonl; done
cat gpurun_out/r02zr_search_call_probe.jsonl; tail -3 gpurun_out/r02zr_probe.err
