set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_mirror.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02zq_pytest_sched.log 2>&1; tail -4 gpurun_out/r02zq_pytest_sched.log
for sb in 1 0 1 0; do
  for L in 20 18 22; do
    B200KF_TREE_SMALL_BLOCKS=$sb timeout 300 python tools/search_probe.py --L $L --reps 300 2>> gpurun_out/r02zq_probe.err | sed "s/^{/{\"small_blocks\": $sb, /" >> gpurun_out/r02zq_search_probe.jsonl
  done
done
cut -c1-120 gpurun_out/r02zq_search_probe.jsonl
