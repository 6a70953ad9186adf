set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02w_pytest.log 2>&1; tail -4 gpurun_out/r02w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; tail -2 gpurun_out/r02w_smoke.log
timeout 900 python bench.py > gpurun_out/r02w_bench.jsonl 2> gpurun_out/r02w_bench.err; echo "bench rc=$?" >> gpurun_out/r02w_bench.err; tail -3 gpurun_out/r02w_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 2 --warmup 1 --passes-per-step 2 --no-cpu-baseline > gpurun_out/r02w_ncu.log 2>&1; tail -2 gpurun_out/r02w_ncu.log | cut -c1-200
