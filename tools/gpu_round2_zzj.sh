set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02zzj_pytest.log 2>&1; tail -3 gpurun_out/r02zzj_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02zzj_smoke.log 2>&1; tail -1 gpurun_out/r02zzj_smoke.log
timeout 900 python bench.py > gpurun_out/r02zzj_bench.jsonl 2> gpurun_out/r02zzj_bench.err; echo "bench rc=$?"
