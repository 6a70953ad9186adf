set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02zze_pytest.log 2>&1; tail -4 gpurun_out/r02zze_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02zze_smoke.log 2>&1; tail -2 gpurun_out/r02zze_smoke.log
timeout 900 python bench.py > gpurun_out/r02zze_bench.jsonl 2> gpurun_out/r02zze_bench.err; echo "bench rc=$?" >> gpurun_out/r02zze_bench.err; tail -2 gpurun_out/r02zze_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02zze_bench_reference_arm.jsonl 2> gpurun_out/r02zze_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02zze_bench_reference_arm.jsonl
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:score_ --csv --log-file gpurun_out/r02zze_search20_fp64.csv python tools/search_probe.py --L 20 --reps 1 > gpurun_out/r02zze_ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:score_ --csv --log-file gpurun_out/r02zze_search26_fp64.csv python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02zze_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tree -s 2 -c 1 -f -o gpurun_out/r02zze_score_tree python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02zze_ncu3.log 2>&1; tail -2 gpurun_out/r02zze_ncu3.log
