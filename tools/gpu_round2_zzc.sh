set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:kf_run_tma_kernelILi0EdLb1ELb0ELi896ELb0E -s 3 -c 1 -f -o gpurun_out/r02zzc_filter_kernel python bench.py --steps 1 --warmup 3 --passes-per-step 3 --no-e2e --no-cpu-baseline --no-search --no-traj --no-shared --no-configs > gpurun_out/r02zzc_ncu.log 2>&1; tail -3 gpurun_out/r02zzc_ncu.log | cut -c1-200
ls -la gpurun_out/r02zzc*
