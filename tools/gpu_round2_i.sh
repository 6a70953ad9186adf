set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense_axis.py tests/test_gpu_run_batch.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02i_pytest.log 2>&1; tail -15 gpurun_out/r02i_pytest.log
timeout 600 python tools/bench_dense.py > gpurun_out/r02i_dense.jsonl 2> gpurun_out/r02i_dense.err; tail -3 gpurun_out/r02i_dense.err; cat gpurun_out/r02i_dense.jsonl
