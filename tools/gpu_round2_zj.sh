set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02zj_pytest.log 2>&1; tail -6 gpurun_out/r02zj_pytest.log
timeout 900 python bench.py > gpurun_out/r02zj_bench.jsonl 2> gpurun_out/r02zj_bench.err; echo "bench rc=$?" >> gpurun_out/r02zj_bench.err; tail -3 gpurun_out/r02zj_bench.err
B200KF_SHARED=0 timeout 400 python tools/bench_models.py --tsteps 1500 --models BIAS9,CV6 2>&1 | grep '"mode": "auto"' > gpurun_out/r02zj_models.jsonl
