set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zf_smoke.log 2>&1; tail -2 gpurun_out/r02zf_smoke.log
timeout 900 python bench.py > gpurun_out/r02zf_bench.jsonl 2> gpurun_out/r02zf_bench.err; echo "bench rc=$?" >> gpurun_out/r02zf_bench.err; tail -3 gpurun_out/r02zf_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02zf_bench_ref.jsonl 2> gpurun_out/r02zf_bench_ref.err; cut -c1-200 gpurun_out/r02zf_bench_ref.jsonl
