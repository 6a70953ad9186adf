set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_small_tiles.py tests/test_gpu_mirror.py tests/test_gpu_schedules.py tests/test_gpu_length_parity.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02q_pytest.log 2>&1; tail -15 gpurun_out/r02q_pytest.log
for st in 1 0 all; do
B200KF_SWEEP_SMALL_TILES=$st timeout 600 python tools/config2_pipe_probe.py > gpurun_out/r02q_pipe_probe_st$st.json 2> gpurun_out/r02q_pipe_probe_st$st.err; echo "probe rc=$?"; tail -3 gpurun_out/r02q_pipe_probe_st$st.err
done
