"""The reference's experiment driver (the body of `kf_workers.py` `__main__`, :2298-2396) on the B200
engine: for a list of windows, greedy (adaptive-threshold) vs optimal (brute-force) vs no-update
log-det, one row per window with the columns `KalmanFilterLogger` writes (:2150-2162).

Per window the reference runs (1) a warm-up adaptive filter from the START of the stream up to the
window (:2316 -- its real wall-clock sink: O(10^5) events x 284 windows), (2) the full filter, (3)
the adaptive filter, (4) the brute-force search and (5) the no-update filter over the window
(:2320-2344).  Here (2)-(5) are the drop-in methods of `kf_workers.KF_SensorFusion` (GPU kernels),
and the warm-ups are NOT repeated from zero: the adaptive filter is causal, so all windows that use
the same threshold share ONE pass over the stream, cut into launches at the window starts with the
full 15-state x and P carried on the device (`adaptive_warmups`).  The numbers are the ones the
reference computes window by window.
"""
from __future__ import annotations

import csv
import os
import random
from datetime import datetime

import numpy as np

from . import streams
from .kf_workers import find_start_idx_for_time_offset

COLUMNS = ["timestamp", "iteration", "start_idx", "num_greedy_measurements", "num_optimal_measurements",
           "initial_optimal_log_det", "final_greedy_log_det", "final_noupdate_log_det", "final_optimal_log_det",
           "r_value", "start_offset"]  # kf_workers.py:2150-2162


class ExperimentLog:
    """CSV sink with the header and row layout of KalmanFilterLogger (kf_workers.py:2143-2196)."""

    def __init__(self, log_file="kf_experiment_results.csv"):
        self.log_file = log_file
        new = not os.path.exists(log_file)
        with open(log_file, "a", newline="") as f:
            if new:
                csv.writer(f).writerow(COLUMNS)

    def log_row(self, row):
        with open(self.log_file, "a", newline="") as f:
            csv.writer(f).writerow([row[c] for c in COLUMNS])


def adaptive_warmups(sf, end_indices, thresholds):
    """For every (end_idx, R_threshold) pair: (states[-1], P) of
    `sf.run_adaptive_threshold_kalman_filter(end_idx=end_idx, R_threshold=R_threshold)` (:2316),
    or None where the reference returns None (no GPS before end_idx).  The adaptive filter is causal, so
    ONE pass over the stream serves every pair: one filter per distinct threshold, all in the same launches
    (b200kf_run_batch_gated: a gate threshold per filter), cut at the requested end indices."""
    ps = streams.pack_events(sf.indexed_sensor_data)
    n_ev = len(sf.indexed_sensor_data)
    out = [None] * len(end_indices)
    if not len(end_indices):
        return out
    thr_list = sorted({float(t) for t in thresholds})
    g_of = {t: g for g, t in enumerate(thr_list)}
    # an unreachable window start (find_start_idx_for_time_offset -> None) is the caller's ERROR row, not a warm-up
    reqs = sorted((min(int(e), n_ev), g_of[float(t)], k) for k, (e, t) in enumerate(zip(end_indices, thresholds))
                  if e is not None)
    if not reqs:
        return out
    e_max = reqs[-1][0]
    init = sf._first_gps(0, e_max)
    if init is None:
        return out
    xt, prev_time, first = init
    Pt = sf._default_P0().astype(np.float64)
    keep, dts, _ = streams.schedule_dts(ps.times[first:e_max], prev_time, skip_advances_prev=False)
    G = len(thr_list)
    gates = np.array([t if t != -float("inf") else -1.0e300 for t in thr_list])
    eng = sf._engine()
    # carried on the device between launches (full 15-state x), one filter per threshold
    x_d, P_d = np.repeat(xt[None], G, axis=0), np.repeat(Pt[None], G, axis=0)
    last_time = prev_time
    done = 0  # events of `keep` already filtered
    for e, g, k in reqs:
        if e <= first:  # the window starts before the first GPS fix: the reference finds no GPS (:995-1004)
            continue
        upto = int(np.searchsorted(keep, e - first, side="left"))
        if upto > done:
            idx = first + keep[done:upto]
            pay = np.repeat(ps.payload(sf.MODEL_NAME, idx)[:, :, None], G, axis=2)
            # only the state and covariance at the window start are used (:2316-2323): no per-step outputs
            r = eng.run_batch(ps.types[idx], dts[done:upto], pay, x_d, P_d, want_agg=False,
                              gate_threshold=gates if G > 1 else float(gates[0]))
            x_d, P_d = r["x"], r["P"]
            last_time = float(ps.times[idx[-1]])
            done = upto
        if isinstance(x_d, np.ndarray):
            out[k] = ((last_time, *x_d[g, :6]), P_d[g].copy())
        else:
            out[k] = ((last_time, *x_d[g, :6].cpu().numpy()), P_d[g].cpu().numpy().copy())
    return out


def run_window_experiment(sf, start_indices, r_factors, *, start_offset=25, lb_end_idx=100000, iterations=None,
                          log=None):
    """kf_workers.py:2301-2358 for the given windows.  r_factors: the factor of `lb_r_value` per
    window (the reference draws it with random.choice from {0.2, ..., 0.8}, :2309).  Returns the rows
    (dicts with COLUMNS, plus `selected_mask`); a window is logged only when the greedy filter used
    at least one measurement (:2349), failures are logged as 'ERROR' rows (:2375-2392)."""
    _, ld_full, _, _ = sf.run_kalman_filter_full(start_idx=0, end_idx=lb_end_idx)
    lb = min(ld_full)
    print("Lower Bounded R-Value: ", lb)
    r0 = [lb * f for f in r_factors]
    warm = adaptive_warmups(sf, start_indices, r0)
    rows = []
    iterations = list(range(len(start_indices))) if iterations is None else list(iterations)
    for it, start_idx, r_value, w in zip(iterations, start_indices, r0, warm):
        row = {"timestamp": datetime.now().isoformat(), "iteration": it, "start_idx": start_idx,
               "start_offset": start_offset}
        try:
            if start_idx is None:  # the offset is past the end of the stream (:2310 returned None): `None + 25` at :2321
                raise TypeError("unsupported operand type(s) for +: 'NoneType' and 'int'")
            if w is None:
                raise TypeError("cannot unpack non-iterable NoneType object")  # what the reference hits at :2316
            st_last, pt = w
            _, ld_w, _, _ = sf.run_kalman_filter_full(start_idx=start_idx, end_idx=start_idx + start_offset,
                                                      initial_pt=pt, initial_state=st_last, print_output=True)
            r_value = (r_value / lb) * min(ld_w)  # :2326
            _, ldg, _, _, mtg = sf.run_adaptive_threshold_kalman_filter(
                start_idx=start_idx, end_idx=start_idx + start_offset, initial_pt=pt, initial_state=st_last,
                R_threshold=r_value, print_output=True)
            bf = sf.run_brute_force_kalman_filter_no_sampling_min_usage(
                start_idx=start_idx, end_idx=start_idx + start_offset, initial_pt=pt, initial_state=st_last,
                R_threshold=r_value)
            _, ldn, _, _, _ = sf.run_no_update_kalman_filter(
                start_idx=start_idx, end_idx=start_idx + start_offset, initial_pt=pt, initial_state=st_last,
                R_threshold=r_value, print_output=True)
            row.update(num_greedy_measurements=len(mtg),
                       num_optimal_measurements=len(bf["selected_sensors"]) if bf else 0,
                       initial_optimal_log_det=bf["log_determinants"][0] if bf else "N/A",
                       final_greedy_log_det=ldg[-1] if ldg else "N/A", final_noupdate_log_det=ldn[-1],
                       final_optimal_log_det=bf["log_determinants"][-1] if bf else "N/A", r_value=r_value,
                       selected_mask=bf["selected_mask"] if bf else None, logged=len(mtg) != 0)
        except Exception as e:  # noqa: BLE001  (the reference catches everything and logs an ERROR row)
            print(f"Iteration {it} failed with error: {e}")
            row.update(num_greedy_measurements=0, num_optimal_measurements=0, initial_optimal_log_det="ERROR",
                       final_greedy_log_det="ERROR", final_noupdate_log_det="ERROR", final_optimal_log_det="ERROR",
                       r_value=r_value, selected_mask=None, logged=True)
        rows.append(row)
        if log is not None and row["logged"]:
            log.log_row(row)
    return rows


def reference_window_plan(sf, *, first_iteration=16, num_off_data_points=300, low_end=25, high_end=2800,
                          r_value_set=(0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8), seed=None):
    """The windows of the reference's own run (:2306-2313): iteration i starts at the first event at
    or after t0 + int(i (high_end - low_end) / num_off_data_points) seconds, with an r factor drawn
    from r_value_set (seeded here; the reference's draw is unseeded).  Returns (iterations,
    start_indices, r_factors) for run_window_experiment."""
    rng = random.Random(seed)
    its = list(range(first_iteration, num_off_data_points))
    starts = [find_start_idx_for_time_offset(sf, int(i * (high_end - low_end) / num_off_data_points)) for i in its]
    factors = [rng.choice(sorted(r_value_set)) for _ in its]
    return its, starts, factors


__all__ = ["COLUMNS", "ExperimentLog", "adaptive_warmups", "run_window_experiment", "reference_window_plan"]
