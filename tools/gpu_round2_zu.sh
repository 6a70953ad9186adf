set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tree -s 2 -c 1 -f -o gpurun_out/r02zu_score_tree python tools/search_probe.py --L 26 --reps 1 > gpurun_out/r02zu_ncu.log 2>&1; tail -2 gpurun_out/r02zu_ncu.log
