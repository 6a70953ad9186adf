set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/search_multi_gpu.py > gpurun_out/r02zzd_search_n2.log 2>&1; tail -4 gpurun_out/r02zzd_search_n2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02zzd_bench_n2.jsonl 2> gpurun_out/r02zzd_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/r02zzd_bench_n2.err
