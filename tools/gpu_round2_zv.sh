set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_mirror.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02zv_pytest_sched.log 2>&1; tail -4 gpurun_out/r02zv_pytest_sched.log
for pdl in 1 0 1 0; do
  for L in 20 16; do
    B200KF_PDL=$pdl timeout 300 python tools/search_call_probe.py --L $L 2>> gpurun_out/r02zv_probe.err | sed "s/^{/{\"pdl\": $pdl, /" >> gpurun_out/r02zv_search_call_probe.jsonl
  done
done
cut -c1-330 gpurun_out/r02zv_search_call_probe.jsonl; tail -3 gpurun_out/r02zv_probe.err
