set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
for v in base s33 base s33; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  for m in CV6 BIAS9 CA15; do
    timeout 300 python tools/traj_probe.py --reps 300 --model $m >> gpurun_out/r02y_traj_$v.jsonl 2>&1
  done
done
unset B200KF_LIB
timeout 900 python -m pytest tests/test_gpu_run_batch.py -m gpu -q --timeout 600 -p no:cacheprovider -k "output_ring or constant_acceleration or cta_size" > gpurun_out/r02y_pytest.log 2>&1; tail -3 gpurun_out/r02y_pytest.log
