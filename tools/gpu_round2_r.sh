set -x
mkdir -p gpurun_out
timeout 600 python tools/config2_pipe_probe.py > gpurun_out/r02r_pipe_probe.json 2> gpurun_out/r02r_pipe_probe.err; echo "probe rc=$?"; tail -3 gpurun_out/r02r_pipe_probe.err
B200KF_SWEEP_SMALL_TILES=0 timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02r_bench_st0.jsonl 2> gpurun_out/r02r_bench_st0.err
B200KF_SWEEP_SMALL_TILES=1 timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02r_bench_st1.jsonl 2> gpurun_out/r02r_bench_st1.err
B200KF_SWEEP_SMALL_TILES=all timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02r_bench_stall.jsonl 2> gpurun_out/r02r_bench_stall.err
