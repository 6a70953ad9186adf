set -x
mkdir -p gpurun_out
timeout 600 python tools/sanitize_search.py > gpurun_out/r02zzf_search_cases.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02zzf_search_cases.log
