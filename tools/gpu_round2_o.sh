set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r02o_bench_n2.jsonl 2> gpurun_out/r02o_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02o_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r02o_bench_ref_n2.jsonl 2> gpurun_out/r02o_bench_ref_n2.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02o_bench_ref_n2.jsonl
