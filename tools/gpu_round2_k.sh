set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_matrix_probe.py > gpurun_out/r02k_h2d_matrix_n4.json 2> gpurun_out/r02k_h2d_matrix_n4.err; echo "probe rc=$?"
tail -c 300 gpurun_out/r02k_h2d_matrix_n4.err
B200KF_SHARED=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:kf_run_tma -s 1 -c 1 -o gpurun_out/r02k_ca15_ring_agg python tools/bench_models.py --models CA15 --tsteps 1000 > gpurun_out/r02k_ncu1.log 2>&1
ls -la gpurun_out | tail -5
