set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02zzb_bench_n8.jsonl 2> gpurun_out/r02zzb_bench_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/r02zzb_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/search_multi_gpu.py > gpurun_out/r02zzb_search_n8.log 2>&1; tail -4 gpurun_out/r02zzb_search_n8.log | cut -c1-300
