set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02zzg_bench_n4.jsonl 2> gpurun_out/r02zzg_bench_n4.err; echo "bench rc=$?"; tail -2 gpurun_out/r02zzg_bench_n4.err
