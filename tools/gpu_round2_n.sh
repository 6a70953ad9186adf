set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02n_pytest.log 2>&1; tail -5 gpurun_out/r02n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; tail -2 gpurun_out/r02n_smoke.log
timeout 900 python bench.py > gpurun_out/r02n_bench.jsonl 2> gpurun_out/r02n_bench.err; echo "bench rc=$?" >> gpurun_out/r02n_bench.err; tail -3 gpurun_out/r02n_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02n_bench_ref.jsonl 2> gpurun_out/r02n_bench_ref.err; tail -2 gpurun_out/r02n_bench_ref.jsonl | cut -c1-400
