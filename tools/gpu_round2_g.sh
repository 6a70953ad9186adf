set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02g_pytest.log 2>&1; tail -6 gpurun_out/r02g_pytest.log
timeout 900 python bench.py > gpurun_out/r02g_bench.jsonl 2> gpurun_out/r02g_bench.err; echo "bench rc=$?" >> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
B200KF_SHARED=0 timeout 600 python tools/bench_models.py --tsteps 1000 > gpurun_out/r02g_models.jsonl 2>&1
