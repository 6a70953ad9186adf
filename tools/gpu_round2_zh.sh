set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
B="--no-cpu-baseline --no-configs --no-search --no-e2e --no-shared --no-traj --steps 12 --warmup 3"
for v in h333 h332 h343 base h333 h332 h343; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  timeout 300 python bench.py $B >> gpurun_out/r02zh_head_$v.jsonl 2>&1
done
export B200KF_LIB=$PWD/$L/libb200kf_h333.so
timeout 1700 python -m pytest tests/test_gpu_run_batch.py tests/test_gpu_full_size.py tests/test_gpu_length_parity.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02zh_pytest_h333.log 2>&1; tail -12 gpurun_out/r02zh_pytest_h333.log
