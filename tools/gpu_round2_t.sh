set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02t_pytest.log 2>&1; tail -15 gpurun_out/r02t_pytest.log
timeout 600 python tools/config2_pipe_probe.py > gpurun_out/r02t_pipe_probe.json 2> gpurun_out/r02t_pipe_probe.err; echo "probe rc=$?"; tail -3 gpurun_out/r02t_pipe_probe.err
timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02t_bench.jsonl 2> gpurun_out/r02t_bench.err
