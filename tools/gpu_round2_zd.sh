set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_run_batch.py tests/test_gpu_schedules.py -m gpu -q --timeout 900 -p no:cacheprovider -k "CVBU6" > gpurun_out/r02zd_pytest.log 2>&1; tail -15 gpurun_out/r02zd_pytest.log
B200KF_SHARED=0 timeout 300 python tools/bench_models.py --models CVBU6,CV6 --tsteps 2000 > gpurun_out/r02zd_models.jsonl 2>&1
timeout 300 python tools/bench_models.py --models CVBU6 --tsteps 2000 > gpurun_out/r02zd_models_shared.jsonl 2>&1
