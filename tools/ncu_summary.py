"""Text summary of one `ncu --set full --import-source on` capture for profiles/ (read HERE, no GPU needed):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "title line" > profiles/x.txt

Prints the launch geometry, DRAM / pipe / issue metrics, the warp-stall ratios and the SASS instructions that
collect the most warp-stall samples."""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")


def page(name, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
h, unit, v = raw[0], raw[1], raw[-1]
print("#", title)
print("# kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
for i, n in enumerate(h):
    if any(n.startswith(k) for k in KEYS) or ("issue_stalled" in n and "per_issue_active" in n):
        print(f"{n:95s} {unit[i]:10s} {v[i]}")
src = page("source", "--print-source", "sass")
hs = src[1]
ix = {n: i for i, n in enumerate(hs)}
data = [r for r in src[2:] if len(r) == len(hs)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(f"\n# warp-stall samples by SASS instruction (top 10 of {tot})")
ST = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_not_selected", "stall_dispatch", "stall_math"]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:10]:
    st = {k: int(r[ix[k]] or 0) for k in ST}
    print(r[ix["Address"]][-5:], r[ix["Source"]][:58].ljust(58), r[ix["# Samples"]].rjust(6),
          {k: c for k, c in st.items() if c})
