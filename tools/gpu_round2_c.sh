set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_schedules.py -m gpu -q --timeout 600 -p no:cacheprovider -k "epsilon or warp_coop" > gpurun_out/r02c_pytest.log 2>&1; tail -3 gpurun_out/r02c_pytest.log
timeout 300 python tools/traj_probe.py --reps 100 > gpurun_out/r02c_traj_probe.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_run_tma -s 3 -c 1 -o gpurun_out/r02c_traj_bulk python tools/traj_probe.py --reps 2 > gpurun_out/r02c_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_greedy_warp -c 1 -o gpurun_out/r02c_select_warp python tools/config2_sweep.py --seconds 30 --seeds 512 > gpurun_out/r02c_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tree -s 2 -c 1 -o gpurun_out/r02c_score_tree python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02c_ncu3.log 2>&1
ls -la gpurun_out | tail -8
