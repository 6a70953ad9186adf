set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 > gpurun_out/r02_bench_n8.jsonl 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n8.err
