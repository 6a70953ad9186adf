set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_schedules.py tests/test_gpu_length_parity.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02l_pytest.log 2>&1; tail -5 gpurun_out/r02l_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02l_bench.jsonl 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02l_bench.err
