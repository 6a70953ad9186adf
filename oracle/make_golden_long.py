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


def main():
    os.makedirs(OUT, exist_ok=True)
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
