set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_run_batch.py -m gpu -q --timeout 600 -p no:cacheprovider -k "constant_acceleration or f64_matches_oracle or f32_mode or gate" > gpurun_out/r02d_pytest.log 2>&1; tail -15 gpurun_out/r02d_pytest.log
timeout 120 python tools/hbm_mix_probe.py > gpurun_out/r02d_hbm_mix.json 2>&1
B200KF_SHARED=0 timeout 600 python tools/bench_models.py --models CA8,CA15 --tsteps 1000 > gpurun_out/r02d_models_ring.jsonl 2>&1
B200KF_SHARED=0 B200KF_CA_RING=0 timeout 600 python tools/bench_models.py --models CA8,CA15 --tsteps 1000 > gpurun_out/r02d_models_ldg.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_run_tma -s 3 -c 1 -o gpurun_out/r02d_ca15_ring python tools/traj_probe.py --model CA15 --reps 2 --mode symmetric > gpurun_out/r02d_ncu1.log 2>&1
ls -la gpurun_out | tail -6
