set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_matrix_probe.py > gpurun_out/r02h_h2d_matrix.json 2> gpurun_out/r02h_h2d_matrix.err; echo "probe rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02h_bench_n8.jsonl 2> gpurun_out/r02h_bench_n8.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02h_h2d_matrix.err
