set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_length_parity.py tests/test_gpu_mirror.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02u_pytest.log 2>&1; tail -8 gpurun_out/r02u_pytest.log
for d in 1 2; do PIPE_DEPTH=$d timeout 600 python tools/config2_pipe_probe.py > gpurun_out/r02u_pipe_probe_d$d.json 2> gpurun_out/r02u_pipe_probe_d$d.err; done
timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02u_bench.jsonl 2> gpurun_out/r02u_bench.err
