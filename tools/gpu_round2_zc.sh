set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kf_run_tma -s 3 -c 1 -o gpurun_out/r02zc_traj_cv6 python tools/traj_probe.py --reps 3 --model CV6 > gpurun_out/r02zc_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kf_run_shared_generic -s 30 -c 1 -o gpurun_out/r02zc_run_gains python tools/config2_pipe_probe.py --seconds 130 > gpurun_out/r02zc_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:select_greedy_warp -s 1 -c 1 -o gpurun_out/r02zc_select python tools/config2_pipe_probe.py --seconds 60 > gpurun_out/r02zc_ncu3.log 2>&1
ls -la gpurun_out | grep r02zc
