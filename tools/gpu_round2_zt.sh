set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tree -s 1 -c 1 -f -o gpurun_out/r02zt_score_tree python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02zt_ncu.log 2>&1; tail -2 gpurun_out/r02zt_ncu.log
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:score_ --csv --log-file gpurun_out/r02zt_search20_fp64.csv python tools/search_probe.py --L 20 --reps 1 > gpurun_out/r02zt_ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:score_ --csv --log-file gpurun_out/r02zt_search26_fp64.csv python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02zt_ncu2.log 2>&1
ls -la gpurun_out | tail -5
