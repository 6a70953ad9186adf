"""TEST INFRASTRUCTURE ONLY -- full-LENGTH golden vectors from the UNMODIFIED reference.

    python oracle/make_golden_long.py            (build container only: needs /root/reference)

tests/golden/*.npz of oracle/make_golden.py pin the path on streams of a few hundred events.  This
script pins it at the lengths BASELINE.json names, where a rounding difference would have had
hundreds of thousands of steps to grow:

* config 2: `run_kalman_filter_scheduled(selection_method='greedy')`
  (kf_workers.py:826-957) over the whole 2 808 s stream (589 680 events, 200 Hz IMU + 10 Hz GPS) at
  the two ends of the sampling sweep (10 Hz and 120 Hz) for four Monte-Carlo noise seeds each;
* config 1 shape: `run_kalman_filter_full` (kf_workers.py:623-728) over the first 120 000 events.

The inputs are NOT stored (42 MB): they are regenerated from `synth.make_indexed_sensor_data` +
`synth.seed_noise` by the tests, and a checksum of them is stored so a drift of the generator is
detected instead of reported as a parity failure.  Outputs stored per run: the selected event
indices (delta-coded, identical for all seeds of a rate), every 499th state and log det plus the
last 8, (max, sum, last) of the log dets, and the final covariance.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.make_golden import arrays_to_events, events_to_arrays, quiet, states_arr  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402
from sensorfusion_kalmanfilter_b200.synth import make_indexed_sensor_data, seed_noise  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SECONDS = 2808.0           # first fix -> end of the reference's data set
N_EVENTS = int(SECONDS * 210)
BASE_SEED = 5
RATES = (10, 120)
SEEDS = (0, 1, 2, 3)
STRIDE = 499
FULL_EVENTS = 120000


def base_arrays(n_events=N_EVENTS):
    ev = make_indexed_sensor_data(n_events + 3, seed=BASE_SEED)[3:]
    assert ev[0][1] == "GPS"
    return events_to_arrays(ev)


def checksum(ty, tm, da):
    return np.array([float(ty.sum()), float(tm.sum()), float(np.abs(da).sum()), float((da * da).sum())])


def sample(a):
    a = np.asarray(a)
    return np.concatenate([a[:-8:STRIDE], a[-8:]], axis=0)


def _sched_job(job):
    rate, seed = job
    kw, _ = load_reference()
    ty, tm, da = base_arrays()
    ev = arrays_to_events(ty, tm, da + seed_noise(ty, seed))
    kf = kw.KF_SensorFusion("gps.csv", "imu.csv")
    kf.indexed_sensor_data = ev
    kf.set_processing_frequency(rate)
    st, ld, P = quiet(kf.run_kalman_filter_scheduled, 0, len(ev), None, None, "greedy")
    st = states_arr(st)
    ld = np.array(ld, dtype=np.float64)
    sel = np.searchsorted(tm, st[1:, 0])
    assert np.array_equal(tm[sel], st[1:, 0])
    return rate, seed, {"n_states": np.int64(len(st)), "sel_delta": np.diff(sel, prepend=0).astype(np.int32),
                        "states": sample(st), "logdet": sample(ld),
                        "agg": np.array([ld.max(), ld.sum(), ld[-1]]), "P": np.asarray(P, dtype=np.float64)}


def _full_job(_):
    kw, _h = load_reference()
    ty, tm, da = base_arrays(FULL_EVENTS)
    ev = arrays_to_events(ty, tm, da)
    kf = kw.KF_SensorFusion("gps.csv", "imu.csv")
    kf.indexed_sensor_data = ev
    # the reference keeps a copy of P per step in _ground_truth_cov (kf_workers.py:723): 1.8 KB x 120 000 = 216 MB
    st, ld, P, prev = quiet(kf.run_kalman_filter_full, 0, len(ev))
    st = states_arr(st)
    ld = np.array(ld, dtype=np.float64)
    return {"checksum": checksum(ty, tm, da), "n_states": np.int64(len(st)), "states": sample(st), "logdet": sample(ld),
            "agg": np.array([ld.max(), ld.sum(), ld[-1]]), "P": np.asarray(P, dtype=np.float64),
            "prev_time": np.float64(prev)}


CFG1_IMU_ROWS = 160000  # 800 s of 200 Hz IMU: the 273 s without a fix + 527 s of fixes (5 270 of them)


def config1():
    """BASELINE configs[0] on the reference's ONE real input (SURVEY.md section 4 data pins): gps_data.csv is
    stored as four float64 columns (tests/golden/gps_data_reference.npz, every field round-trips through
    repr, so synth.write_gps_csv rebuilds the file byte for byte -- sha256 stored) and the UNMODIFIED
    front ends + filters of hw5_2.py and kf_workers.py are run on it with a synthetic imu_data.csv
    (synth.write_config1_csvs, hw5_1.py:29-31 schema).  `utm.from_latlon` is the product's restatement
    (geo.py), as in make_golden.py csv: the UTM series itself stays unpinned."""
    import csv
    import hashlib
    import tempfile
    from oracle.ref_import import REFERENCE_DIR
    from sensorfusion_kalmanfilter_b200 import geo, synth
    kw, hw = load_reference()
    sys.modules["utm"].from_latlon = geo.from_latlon
    src = os.path.join(REFERENCE_DIR, "gps_data.csv")
    raw = open(src, "rb").read()
    rows = list(csv.reader(open(src, newline="")))[1:]
    cols = np.array([[float(v) for v in r] for r in rows], dtype=np.float64)
    g = {"sha256": np.frombuffer(hashlib.sha256(raw).digest(), dtype=np.uint8), "n_bytes": np.int64(len(raw)),
         "time": cols[:, 0], "lat": cols[:, 1], "lon": cols[:, 2], "alt": cols[:, 3]}
    np.savez_compressed(os.path.join(OUT, "gps_data_reference.npz"), **g)
    out = {"imu_rows": np.int64(CFG1_IMU_ROWS), "stride": np.int64(STRIDE)}
    with tempfile.TemporaryDirectory() as d:
        gps_path = synth.write_gps_csv(os.path.join(d, "gps_data.csv"), *[cols[:, j] for j in range(4)])
        assert open(gps_path, "rb").read() == raw
        _, imu_path = synth.write_config1_csvs(d, n_imu=CFG1_IMU_ROWS, seed=0, gps_csv=gps_path)
        for tag, mod in (("kw", kw), ("hw", hw)):
            k = mod.KF_SensorFusion(gps_path, imu_path)
            quiet(k.load_data)
            quiet(k.gps_to_modified_utm if tag == "kw" else k.gps_to_utm)
            wb, ab, fvi = quiet(k.compute_imu_biases, k.gps_data, k.imu_data)
            quiet(k.unbias_imu_data, wb, ab)
            quiet(k.combine_sensor_data)
            ev = k.indexed_sensor_data
            out.update({f"{tag}_n_events": np.int64(len(ev)), f"{tag}_n_fixes": np.int64(len(k.utm_data)),
                        f"{tag}_first_valid": np.int64(fvi), f"{tag}_wbias": wb, f"{tag}_abias": ab,
                        f"{tag}_first_fix_time": np.float64(k.utm_data[0]["time"]),
                        f"{tag}_zone": np.array([k.utm_data[0]["zone_number"], ord(k.utm_data[0]["zone_letter"])])})
            if tag == "kw":
                out["kw_first_fix_alt"] = np.float64(k.utm_data[0]["altitude"])
                st, ld, P, prev = quiet(k.run_kalman_filter_full, 0, len(ev))
                st, ld = states_arr(st), np.array(ld, dtype=np.float64)
                out.update(kw_n_states=np.int64(len(st)), kw_states=sample(st), kw_logdet=sample(ld),
                           kw_agg=np.array([ld.max(), ld.sum(), ld[-1]]), kw_P=np.asarray(P, dtype=np.float64))
            else:
                st = np.array(quiet(k.run_kalman_filter), dtype=np.float64)
                out.update(hw_n_states=np.int64(len(st)), hw_states=sample(st))
    np.savez_compressed(os.path.join(OUT, "config1_real_gps.npz"), **out)
    for f in ("gps_data_reference.npz", "config1_real_gps.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        config1()
        return
    ctx = mp.get_context("fork")
    jobs = [(r, s) for r in RATES for s in SEEDS]
    with ctx.Pool(min(len(jobs) + 1, os.cpu_count() or 2)) as pool:
        full = pool.apply_async(_full_job, (0,))
        res = pool.map(_sched_job, jobs, chunksize=1)
        full = full.get()
    ty, tm, da = base_arrays()
    g = {"n_events": np.int64(N_EVENTS), "base_seed": np.int64(BASE_SEED), "rates": np.array(RATES),
         "seeds": np.array(SEEDS), "stride": np.int64(STRIDE), "checksum": checksum(ty, tm, da),
         "noise_checksum": np.array([float(np.abs(seed_noise(ty, s)).sum()) for s in SEEDS])}
    for rate, seed, out in res:
        for k, v in out.items():
            if k == "sel_delta":
                if seed == SEEDS[0]:
                    g[f"r{rate}_sel_delta"] = v
                else:  # the greedy selection never reads a measurement (kf_workers.py:174-213)
                    assert np.array_equal(v, g[f"r{rate}_sel_delta"])
            else:
                g[f"r{rate}_s{seed}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ca15_config2_full_length.npz"), **g)
    np.savez_compressed(os.path.join(OUT, "ca15_full_120k.npz"), n_events=np.int64(FULL_EVENTS),
                        base_seed=np.int64(BASE_SEED), stride=np.int64(STRIDE), **full)
    for f in ("ca15_config2_full_length.npz", "ca15_full_120k.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
