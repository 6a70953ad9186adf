set -x
mkdir -p gpurun_out
timeout 300 python tools/search_shard_probe.py --L 32 --world 8 > gpurun_out/r02zl_shard_probe.json 2> gpurun_out/r02zl_shard_probe.err; tail -2 gpurun_out/r02zl_shard_probe.err; cat gpurun_out/r02zl_shard_probe.json
timeout 300 python tools/search_shard_probe.py --L 32 --world 1 >> gpurun_out/r02zl_shard_probe.json 2>> gpurun_out/r02zl_shard_probe.err
