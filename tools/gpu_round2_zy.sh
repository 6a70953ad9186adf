set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02zy_pytest.log 2>&1; tail -4 gpurun_out/r02zy_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02zy_smoke.log 2>&1; tail -2 gpurun_out/r02zy_smoke.log
( time timeout 900 python bench.py > gpurun_out/r02zy_bench.jsonl 2> gpurun_out/r02zy_bench.err ) 2> gpurun_out/r02zy_bench_time.txt; echo "bench rc=$?" >> gpurun_out/r02zy_bench.err; tail -3 gpurun_out/r02zy_bench.err; cat gpurun_out/r02zy_bench_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02zy_launches.csv python bench.py --steps 2 --warmup 1 --passes-per-step 3 --no-cpu-baseline --no-configs > gpurun_out/r02zy_ncu_launches.log 2>&1; tail -2 gpurun_out/r02zy_ncu_launches.log | cut -c1-200
