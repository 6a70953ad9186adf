set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_length_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "warp_coop or chunked_selection or pipelined or chunked_sweep" > gpurun_out/r02f_pytest.log 2>&1; tail -8 gpurun_out/r02f_pytest.log
timeout 900 python bench.py > gpurun_out/r02f_bench.jsonl 2> gpurun_out/r02f_bench.err; echo "bench rc=$?" >> gpurun_out/r02f_bench.err; tail -5 gpurun_out/r02f_bench.err
