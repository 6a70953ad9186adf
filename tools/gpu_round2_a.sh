set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 900 python bench.py > gpurun_out/r02a_bench.jsonl 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 300 python tools/search_probe.py --L 20 > gpurun_out/r02a_search20.json 2>&1
timeout 300 python tools/search_probe.py --L 26 > gpurun_out/r02a_search26.json 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:score_ --csv --log-file gpurun_out/r02a_search20_fp64.csv python tools/search_probe.py --L 20 --reps 1 > gpurun_out/r02a_ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:score_ --csv --log-file gpurun_out/r02a_search26_fp64.csv python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02a_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tree -s 1 -c 1 -o gpurun_out/r02a_score_tree python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02a_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_greedy -c 1 -o gpurun_out/r02a_select_greedy python tools/config2_sweep.py --seconds 30 --seeds 512 > gpurun_out/r02a_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 2 --warmup 1 --passes-per-step 3 --no-cpu-baseline --no-configs > gpurun_out/r02a_ncu5.log 2>&1
ls -la gpurun_out
