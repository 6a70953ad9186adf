set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_mirror.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02zn_pytest_sched.log 2>&1; tail -4 gpurun_out/r02zn_pytest_sched.log
timeout 300 python tools/search_shard_probe.py --L 32 --world 8 > gpurun_out/r02zn_shard_probe.json 2> gpurun_out/r02zn_shard_probe.err; tail -2 gpurun_out/r02zn_shard_probe.err; cat gpurun_out/r02zn_shard_probe.json
timeout 300 python tools/search_shard_probe.py --L 32 --world 8 --contiguous >> gpurun_out/r02zn_shard_probe.json 2>> gpurun_out/r02zn_shard_probe.err
timeout 300 python tools/search_shard_probe.py --L 26 --world 8 >> gpurun_out/r02zn_shard_probe.json 2>> gpurun_out/r02zn_shard_probe.err
timeout 300 python tools/search_shard_probe.py --L 26 --world 8 --contiguous >> gpurun_out/r02zn_shard_probe.json 2>> gpurun_out/r02zn_shard_probe.err
tail -3 gpurun_out/r02zn_shard_probe.json | cut -c1-900
