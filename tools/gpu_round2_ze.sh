set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02ze_pytest.log 2>&1; tail -15 gpurun_out/r02ze_pytest.log
B200KF_SHARED=0 timeout 300 python tools/bench_models.py --models CVBU6 --tsteps 2000 > gpurun_out/r02ze_models.jsonl 2>&1
