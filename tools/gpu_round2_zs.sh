set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02zs_pytest.log 2>&1; tail -6 gpurun_out/r02zs_pytest.log
