set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
for v in base h333; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  B200KF_SHARED=0 timeout 400 python tools/bench_models.py --tsteps 1500 2>&1 | grep '"mode": "auto"' > gpurun_out/r02zi_models_$v.jsonl
  timeout 400 python tools/bench_models.py --tsteps 1500 2>&1 | grep '"mode": "auto"' > gpurun_out/r02zi_models_shared_$v.jsonl
  B200KF_SHARED=0 timeout 400 python tools/bench_models.py --tsteps 3000 --filters 32768 --models CV6,BIAS9 2>&1 | grep '"mode": "auto"' > gpurun_out/r02zi_models_32k_$v.jsonl
done
