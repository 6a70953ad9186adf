set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
B="--no-cpu-baseline --no-configs --no-search --no-e2e --no-shared --no-traj --steps 12 --warmup 3"
for v in base h252 h243 base h252 h243; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  timeout 300 python bench.py $B >> gpurun_out/r02zk_head_$v.jsonl 2>&1
done
