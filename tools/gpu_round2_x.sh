set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r02x_bench_n2.jsonl 2> gpurun_out/r02x_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02x_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/search_multi_gpu.py > gpurun_out/r02x_search_n2.log 2>&1; tail -3 gpurun_out/r02x_search_n2.log | cut -c1-300
