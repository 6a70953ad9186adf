set -x
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_matrix_probe.py > gpurun_out/r02h_h2d_matrix_n8.json 2> gpurun_out/r02h_h2d_matrix_n8.err; echo "probe rc=$?"
tail -c 300 gpurun_out/r02h_h2d_matrix_n8.err
