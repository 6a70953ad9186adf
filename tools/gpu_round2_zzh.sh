set -x
mkdir -p gpurun_out
timeout 300 python tools/host_register_probe.py > gpurun_out/r02zzh_host_register_probe.json 2> gpurun_out/r02zzh.err; cat gpurun_out/r02zzh_host_register_probe.json; tail -3 gpurun_out/r02zzh.err
