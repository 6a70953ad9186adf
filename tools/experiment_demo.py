"""The reference's whole experiment (kf_workers.py:2298-2358: 284 windows, each with a warm-up
adaptive filter from the start of the stream, a 25-event window filtered fully / greedily / without
updates and searched exhaustively over its 2^25 subsets) on a synthetic stream of the data set's
size, timed.

    python tools/experiment_demo.py [--events 590000] [--windows 284] [--offset 25]"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import synth  # noqa: E402
from sensorfusion_kalmanfilter_b200.experiment import reference_window_plan, run_window_experiment  # noqa: E402
from sensorfusion_kalmanfilter_b200.kf_workers import KF_SensorFusion  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=int, default=590000)
ap.add_argument("--windows", type=int, default=284)
ap.add_argument("--offset", type=int, default=25)
a = ap.parse_args()
t0 = time.perf_counter()
sf = KF_SensorFusion("gps_data.csv", "imu_data.csv")
sf.indexed_sensor_data = synth.make_indexed_sensor_data(a.events, seed=9)
t_stream = time.perf_counter() - t0
its, starts, factors = reference_window_plan(sf, first_iteration=300 - a.windows, seed=0)
t0 = time.perf_counter()
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    rows = run_window_experiment(sf, starts, factors, start_offset=a.offset, iterations=its)
dt = time.perf_counter() - t0
fails = [ln for ln in buf.getvalue().splitlines() if "failed with error" in ln]
for ln in fails[:3]:
    print(ln, file=sys.stderr)
ok = [r for r in rows if r["final_greedy_log_det"] != "ERROR"]
found = [r for r in ok if r["selected_mask"] is not None]
print(json.dumps({"events": a.events, "windows": len(rows), "window_events": a.offset,
                  "subsets_per_window": 2 ** a.offset, "stream_build_s": round(t_stream, 2), "experiment_s": round(dt, 2),
                  "windows_ok": len(ok), "windows_with_feasible_schedule": len(found),
                  "mean_greedy_measurements": sum(r["num_greedy_measurements"] for r in ok) / max(1, len(ok)),
                  "mean_optimal_measurements": sum(r["num_optimal_measurements"] for r in found) / max(1, len(found))}))
