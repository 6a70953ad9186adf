set -x
mkdir -p gpurun_out
L=sensorfusion-kalmanfilter_b200
B="--no-cpu-baseline --no-configs --no-search --no-e2e --no-shared --no-traj --steps 10 --warmup 5"
timeout 300 python bench.py $B > gpurun_out/r02e_head_new.jsonl 2>&1
B200KF_LIB=$PWD/$L/libb200kf_aggold.so timeout 300 python bench.py $B > gpurun_out/r02e_head_aggold.jsonl 2>&1
timeout 300 python bench.py $B > gpurun_out/r02e_head_new2.jsonl 2>&1
for v in base s33 evict; do
  if [ $v = base ]; then unset B200KF_LIB; else export B200KF_LIB=$PWD/$L/libb200kf_$v.so; fi
  timeout 300 python tools/traj_probe.py --reps 300 > gpurun_out/r02e_traj_$v.json 2>&1
  timeout 300 python tools/traj_probe.py --reps 300 --model BIAS9 > gpurun_out/r02e_traj_bias9_$v.json 2>&1
done
unset B200KF_LIB
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02e_pytest.log 2>&1; tail -5 gpurun_out/r02e_pytest.log
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
B200KF_SHARED=0 timeout 600 ncu --metrics $M --clock-control none -k regex:kf_run_tma --csv --log-file gpurun_out/r02e_ca_counts.csv python tools/bench_models.py --models CA8,CA15 --tsteps 300 > gpurun_out/r02e_ncu1.log 2>&1
ls gpurun_out | tail -12
