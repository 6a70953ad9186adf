set -x
mkdir -p gpurun_out
for ch in 25200 12600 6300; do
B200KF_CONFIG2_PIPE_CHUNK=$ch timeout 900 python bench.py --no-cpu-baseline --no-search --no-e2e --no-shared --no-traj --steps 3 --warmup 3 > gpurun_out/r02v_bench_ch$ch.jsonl 2> gpurun_out/r02v_bench_ch$ch.err
done
