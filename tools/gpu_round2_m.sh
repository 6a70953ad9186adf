set -x
mkdir -p gpurun_out
timeout 600 python tools/config2_pipe_probe.py > gpurun_out/r02m_pipe_probe.json 2> gpurun_out/r02m_pipe_probe.err; echo "probe rc=$?"; tail -3 gpurun_out/r02m_pipe_probe.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/r02m_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02m_memcheck.log
