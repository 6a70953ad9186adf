set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02j_pytest.log 2>&1; tail -5 gpurun_out/r02j_pytest.log
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:kf_run_axis --csv --log-file gpurun_out/r02j_dense_counts.csv python tools/bench_dense.py --tsteps 100 --reps 1 --precisions f64 --axis 1 > gpurun_out/r02j_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kf_run_axis -s 1 -c 1 -o gpurun_out/r02j_axis_dense python tools/bench_dense.py --tsteps 100 --reps 1 --precisions f64 --axis 1 > gpurun_out/r02j_ncu2.log 2>&1
B200KF_SHARED=0 timeout 600 ncu --metrics $M --clock-control none -k regex:kf_run_tma --csv --log-file gpurun_out/r02j_ca_counts.csv python tools/bench_models.py --models CA8,CA15 --tsteps 300 > gpurun_out/r02j_ncu3.log 2>&1
timeout 900 python bench.py > gpurun_out/r02j_bench.jsonl 2> gpurun_out/r02j_bench.err; echo "bench rc=$?" >> gpurun_out/r02j_bench.err; tail -3 gpurun_out/r02j_bench.err
ls -la gpurun_out | tail -8
