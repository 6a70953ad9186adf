set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -v --timeout 900 -p no:cacheprovider > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
timeout 900 python bench.py > gpurun_out/r02b_bench.jsonl 2> gpurun_out/r02b_bench.err; echo "bench rc=$?" >> gpurun_out/r02b_bench.err
B200KF_BULK_OUT=0 timeout 600 python bench.py --no-cpu-baseline --no-configs --no-search --no-e2e --no-shared > gpurun_out/r02b_bench_nobulk.jsonl 2>&1
ls -la gpurun_out
