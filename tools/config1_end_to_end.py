"""BASELINE configs[0] end to end: gps_data.csv + a synthetic imu_data.csv (hw5_1.py:28-31 schema) ->
the reference's CSV front end -> the 15-state filter of kf_workers.py and the 8-state filter of hw5_2.py
on the GPU (the CPU port of the same single-filter loop is timed by bench.py's cpu_baseline leg:
`single_filter_ca15`).

    python tools/config1_end_to_end.py [--gps-csv /path/to/gps_data.csv] [--imu-rows 60000]

A single filter is the reference's own shape (configs[0] is its CPU-runnable case): one GPU thread
walks the events, so this measures launch + per-step latency, not throughput -- the throughput
configurations are the batched ones bench.py measures.  Prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sensorfusion_kalmanfilter_b200 import hw5_2, kf_workers, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gps-csv", default=None, help="the reference's gps_data.csv (default: a synthetic one)")
ap.add_argument("--imu-rows", type=int, default=60000)
a = ap.parse_args()

with tempfile.TemporaryDirectory() as d:
    gps_path, imu_path = synth.write_config1_csvs(d, n_imu=a.imu_rows, seed=0, gps_csv=a.gps_csv)
    out = {"gps_csv": "reference" if a.gps_csv else "synthetic", "imu_rows": a.imu_rows}
    for name, mod in (("kf_workers_15_state", kf_workers), ("hw5_2_8_state", hw5_2)):
        t0 = time.perf_counter()
        k = mod.KF_SensorFusion(gps_path, imu_path)
        k.load_data()
        k.gps_to_modified_utm() if mod is kf_workers else k.gps_to_utm()
        wb, ab, _ = k.compute_imu_biases(k.gps_data, k.imu_data)
        k.unbias_imu_data(wb, ab)
        k.combine_sensor_data()
        t_front = time.perf_counter() - t0
        n_ev = len(k.indexed_sensor_data)
        run = (lambda: k.run_kalman_filter_full(0, n_ev)) if mod is kf_workers else k.run_kalman_filter
        run()  # warm-up (library load, packing cache)
        t0 = time.perf_counter()
        res = run()
        t_gpu = time.perf_counter() - t0
        n_steps = len(res[0]) - 1 if mod is kf_workers else len(res)
        out[name] = {"events": n_ev, "filter_steps": n_steps, "front_end_s": round(t_front, 3),
                     "gpu_s": round(t_gpu, 4), "gpu_steps_per_s": n_steps / t_gpu}
print(json.dumps(out))
