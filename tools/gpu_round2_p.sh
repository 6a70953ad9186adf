set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_schedules.py tests/test_gpu_length_parity.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02p_pytest.log 2>&1; tail -5 gpurun_out/r02p_pytest.log
timeout 600 python tools/config2_pipe_probe.py > gpurun_out/r02p_pipe_probe.json 2> gpurun_out/r02p_pipe_probe.err; echo "probe rc=$?"; tail -3 gpurun_out/r02p_pipe_probe.err; cut -c1-420 gpurun_out/r02p_pipe_probe.json
