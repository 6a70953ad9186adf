set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
for v in base ca32 base ca32; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  B200KF_SHARED=0 timeout 300 python tools/bench_models.py --models CA15,CA8 --tsteps 1500 2>&1 | grep '"mode": "auto"' | grep '"precision": "f64"' >> gpurun_out/r02zb_models_$v.jsonl
done
