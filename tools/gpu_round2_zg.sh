set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
B="--no-cpu-baseline --no-configs --no-search --no-e2e --no-shared --no-traj --steps 12 --warmup 3"
for v in base h333 h442 base h333 h442; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  timeout 300 python bench.py $B >> gpurun_out/r02zg_head_$v.jsonl 2>&1
done
for v in base ca5 base ca5; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  for m in CA15 CA8; do
    timeout 300 python tools/traj_probe.py --reps 300 --model $m >> gpurun_out/r02zg_traj_$v.jsonl 2>&1
  done
done
unset B200KF_LIB
B200KF_LIB=$PWD/$L/libb200kf_ca5.so timeout 900 python -m pytest tests/test_gpu_run_batch.py -m gpu -q --timeout 600 -p no:cacheprovider -k "output_ring or constant_acceleration" > gpurun_out/r02zg_pytest_ca5.log 2>&1; tail -3 gpurun_out/r02zg_pytest_ca5.log
