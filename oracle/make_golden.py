"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run here (the build container, where /root/reference exists):

    python oracle/make_golden.py

It imports the reference through oracle/ref_import.py, drives its own functions with seeded
synthetic `indexed_sensor_data` streams and stores inputs + outputs as small .npz fixtures.  The
fixtures travel to the GPU box; the reference does not.  tests/test_oracle_golden.py pins
oracle/kf_oracle.py to them, and the `-m gpu` parity tests compare the CUDA path with them.
"""
from __future__ import annotations

import contextlib
import io
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference  # noqa: E402
from sensorfusion_kalmanfilter_b200.synth import make_indexed_sensor_data  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def events_to_arrays(events):
    """Pack reference-layout events as (type u8 [E], time f64 [E], data f64 [E, 9])."""
    ty = np.zeros(len(events), dtype=np.uint8)
    tm = np.zeros(len(events))
    da = np.zeros((len(events), 9))
    for k, (_, stype, t, sdata) in enumerate(events):
        tm[k] = t
        if stype == "GPS":
            ty[k] = 0
            da[k, :3] = [sdata["easting"], sdata["northing"], sdata["altitude"]]
        else:
            ty[k] = 1
            da[k] = [float(v) for v in sdata[1:10]]
    return ty, tm, da


def arrays_to_events(ty, tm, da):
    """Inverse of events_to_arrays (used by the tests to rebuild the exact input)."""
    ev = []
    for k in range(len(ty)):
        t = float(tm[k])
        if ty[k] == 0:
            ev.append((k, "GPS", t, {"time": t, "easting": np.float64(da[k, 0]),
                                     "northing": np.float64(da[k, 1]), "zone_number": 19,
                                     "zone_letter": "T", "altitude": float(da[k, 2])}))
        else:
            ev.append((k, "IMU", t, [repr(t)] + [float(v) for v in da[k]]))
    return ev


def states_arr(states):
    return np.array([[float(v) for v in s] for s in states], dtype=np.float64)


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def notebook_bruteforce():
    """Product-of-queues brute force with the RMSE objective, which exists only in the notebook
    (KF_SensorFusion.ipynb:804-954): cells 0-2 exec'd unmodified (oracle/ref_import.py)."""
    from oracle.ref_import import load_notebook_class
    K = load_notebook_class()
    events = make_indexed_sensor_data(80, seed=31)
    ty, tm, da = events_to_arrays(events)
    events = arrays_to_events(ty, tm, da)
    k = K("gps.csv", "imu.csv")
    k.indexed_sensor_data = events
    k.utm_data = [e[3] for e in events if e[1] == "GPS"]
    g = {"ev_type": ty, "ev_time": tm, "ev_data": da}
    for tag, f, s0, s1 in (("a", 50, 0, 40), ("b", 70, 5, 44)):
        k.set_processing_frequency(f)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            out = k.run_brute_force_kalman_filter(s0, s1)
        g[f"{tag}_freq"] = np.float64(f)
        g[f"{tag}_range"] = np.array([s0, s1])
        g[f"{tag}_selected"] = np.array([e[0] for e in out["selected_sensors"]])
        g[f"{tag}_metric"] = np.float64(out["accuracy_metric"])
        g[f"{tag}_final_state"] = out["final_state"]
        g[f"{tag}_final_cov"] = out["final_covariance"]
        g[f"{tag}_trajectory"] = states_arr(out["trajectory"])
        g[f"{tag}_rmse_of_best"] = np.float64(k.calculate_accuracy_metrics(out["trajectory"]))
    np.savez_compressed(os.path.join(OUT, "nb_bruteforce.npz"), **g)


def accuracy_metrics():
    """kf_workers.calculate_accuracy_metrics (kf_workers.py:1162-1216) of a 30 Hz greedy-scheduled
    trajectory against the full-rate run of the same stream (the reference's __main__ comparison),
    plus the < 2-sample fallback branch (:1185-1187)."""
    kw, _ = load_reference()
    events = make_indexed_sensor_data(420, seed=11)
    ty, tm, da = events_to_arrays(events)
    events = arrays_to_events(ty, tm, da)
    kf = kw.KF_SensorFusion("gps.csv", "imu.csv")
    kf.indexed_sensor_data = events
    quiet(kf.run_kalman_filter_full, 0, 400)  # sets _ground_truth
    kf.set_processing_frequency(30)
    st, _, _ = quiet(kf.run_kalman_filter_scheduled, 40, 360, None, None, "greedy")
    m = kf.calculate_accuracy_metrics(st)
    g = {"ev_type": ty, "ev_time": tm, "ev_data": da, "cand_states": states_arr(st),
         "rmse": np.float64(m["total_position_rmse"]), "position_errors": m["position_errors"],
         "euclidean_errors": m["euclidean_errors"], "ground_truth_interp": m["ground_truth_interp"]}
    tiny = [(events[100][2] + 1e-4, *st[5][1:]), (events[100][2] + 2e-4, *st[6][1:])]  # no GT sample inside
    m2 = kf.calculate_accuracy_metrics(tiny)
    g.update(tiny_states=states_arr(tiny), tiny_rmse=np.float64(m2["total_position_rmse"]),
             tiny_interp=m2["ground_truth_interp"])
    np.savez_compressed(os.path.join(OUT, "ca15_accuracy.npz"), **g)


def experiment_windows():
    """The reference's experiment loop (kf_workers.py:2298-2396, the body of its __main__, which is
    not importable as a function): per window a warm-up adaptive filter from the start of the stream,
    the full / greedy (adaptive) / optimal (brute force) / no-update runs over the window, and the
    row the KalmanFilterLogger would write (:2150-2162).  Every call is the unmodified reference's;
    only the loop that strings them together is restated here, with fixed r factors in place of
    random.choice (:2309) and 8-event windows (2^8 subsets) so that the Python brute force ends."""
    from multiprocessing.pool import ThreadPool
    kw, _ = load_reference()
    kw.Pool = lambda processes=None: ThreadPool(2)  # see the note in main(): forked Pool(30) hangs on early return
    events = make_indexed_sensor_data(620, seed=17)
    ty, tm, da = events_to_arrays(events)
    events = arrays_to_events(ty, tm, da)
    kf = kw.KF_SensorFusion("gps.csv", "imu.csv")
    kf.indexed_sensor_data = events
    lb_end, off = 100, 8
    starts = [130, 275, 410, 505]
    factors = [0.87, 0.94, 0.3, 0.9]  # mid-range thresholds select a few events; 0.3 (the reference's set) selects one
    _, ld_full, _, _ = quiet(kf.run_kalman_filter_full, 0, lb_end)
    lb = min(ld_full)
    rows = []
    for start_idx, fac in zip(starts, factors):
        r_value = lb * fac
        st_g, _, pt_g, _, _ = quiet(kf.run_adaptive_threshold_kalman_filter, end_idx=start_idx, R_threshold=r_value)
        _, ld_w, _, _ = quiet(kf.run_kalman_filter_full, start_idx, start_idx + off, pt_g, st_g[-1], True)
        r_value = (r_value / lb) * min(ld_w)
        _, ldg, _, _, mtg = quiet(kf.run_adaptive_threshold_kalman_filter, start_idx, start_idx + off, r_value,
                                  pt_g, st_g[-1], True)
        bf = quiet(kf.run_brute_force_kalman_filter_no_sampling_min_usage, start_idx, start_idx + off, r_value,
                   pt_g, st_g[-1])
        _, ldn, _, _, _ = quiet(kf.run_no_update_kalman_filter, start_idx, start_idx + off, r_value, pt_g, st_g[-1], True)
        rows.append([start_idx, len(mtg), len(bf["selected_sensors"]) if bf else 0,
                     bf["log_determinants"][0] if bf else np.nan, ldg[-1], ldn[-1],
                     bf["log_determinants"][-1] if bf else np.nan, r_value, off,
                     sum(1 << (e[0] - start_idx) for e in bf["selected_sensors"]) if bf else -1])
    np.savez_compressed(os.path.join(OUT, "ca15_experiment.npz"), ev_type=ty, ev_time=tm, ev_data=da,
                        lb_end=np.int64(lb_end), lb=np.float64(lb), starts=np.array(starts), factors=np.array(factors),
                        rows=np.array(rows, dtype=np.float64))


def csv_front_end():
    """The reference's CSV front end (kf_workers.py:290-385 and hw5_2.py:17-112: load_data ->
    gps_to_modified_utm -> compute_imu_biases -> unbias_imu_data -> combine_sensor_data) run
    unmodified on CSVs written by synth.write_config1_csvs(seed=3), and the first filter states on the
    result.  Harness accommodation: the un-vendored third-party `utm` module the reference imports is
    given the product's restatement (geo.from_latlon) as `utm.from_latlon`, so this pins everything of
    the front end EXCEPT the lat/lon -> UTM series itself (which stays unpinned)."""
    import tempfile
    from sensorfusion_kalmanfilter_b200 import geo, synth
    kw, hw = load_reference()
    sys.modules["utm"].from_latlon = geo.from_latlon
    g = {}
    with tempfile.TemporaryDirectory() as d:
        gps_path, imu_path = synth.write_config1_csvs(d, n_imu=1200, n_gps_nan=8, seed=3)
        for tag, mod in (("kw", kw), ("hw", hw)):
            k = mod.KF_SensorFusion(gps_path, imu_path)
            quiet(k.load_data)
            quiet(k.gps_to_modified_utm if tag == "kw" else k.gps_to_utm)
            wb, ab, fvi = quiet(k.compute_imu_biases, k.gps_data, k.imu_data)
            quiet(k.unbias_imu_data, wb, ab)
            quiet(k.combine_sensor_data)
            ev = k.indexed_sensor_data
            ty = np.array([0 if e[1] == "GPS" else 1 for e in ev], dtype=np.uint8)
            tm = np.array([e[2] for e in ev])
            da = np.zeros((len(ev), 9))
            for i, e in enumerate(ev):
                if e[1] == "GPS":
                    da[i, :2] = [e[3]["easting"], e[3]["northing"]]
                    da[i, 2] = e[3].get("altitude", 0.0)
                else:
                    da[i] = [float(v) for v in e[3][1:10]]
            g.update({f"{tag}_type": ty, f"{tag}_time": tm, f"{tag}_data": da, f"{tag}_wbias": wb, f"{tag}_abias": ab,
                      f"{tag}_first_valid": np.int64(fvi)})
            if tag == "kw":
                st, ld, P, _ = quiet(k.run_kalman_filter_full, 0, 300)
                g.update(kw_states=states_arr(st), kw_logdet=np.array(ld), kw_P=P)
            else:
                st = quiet(k.run_kalman_filter)
                g.update(hw_states=np.array(st, dtype=np.float64))
    np.savez_compressed(os.path.join(OUT, "csv_front_end.npz"), **g)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "csv":
        csv_front_end()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "exp":
        experiment_windows()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "nb":
        notebook_bruteforce()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "acc":
        accuracy_metrics()
        return
    kw, hw = load_reference()

    # ---------------- 15-state filter loops (kf_workers.py) ----------------
    events = make_indexed_sensor_data(420, seed=11)
    # make one out-of-order event to exercise the dt<0 branches (kf_workers.py:683-685,1013-1015)
    ev = list(events)
    k = 130
    ev[k] = (ev[k][0], ev[k][1], ev[k - 1][2] - 0.002, ev[k][3])
    events = ev
    ty, tm, da = events_to_arrays(events)
    events = arrays_to_events(ty, tm, da)  # canonical form that the tests rebuild bit-exactly

    kf = kw.KF_SensorFusion("gps.csv", "imu.csv")
    kf.indexed_sensor_data = events
    g = {"ev_type": ty, "ev_time": tm, "ev_data": da}

    st, ld, P, prev = quiet(kf.run_kalman_filter_full, 0, 400)
    g.update(full_cold_states=states_arr(st), full_cold_logdet=np.array(ld), full_cold_P=P,
             full_cold_prev=np.float64(prev))

    st1, ld1, P1, prev1 = quiet(kf.run_kalman_filter_full, 0, 200)
    init_state = st1[-1]
    st2, ld2, P2, prev2 = quiet(kf.run_kalman_filter_full, 200, 400, P1.copy(), init_state)
    g.update(full_warm_init_state=np.array([float(v) for v in init_state]), full_warm_init_P=P1,
             full_warm_states=states_arr(st2), full_warm_logdet=np.array(ld2), full_warm_P=P2,
             full_warm_prev=np.float64(prev2))

    st, covs = quiet(kf.run_kalman_filter, 0, 300)
    g.update(legacy_states=states_arr(st), legacy_P_last=covs[-1], legacy_P_mid=covs[len(covs) // 2],
             legacy_n=np.int64(len(covs)))

    thr = float(np.median(ld[50:]))
    st, ldl, P, prev, mt = quiet(kf.run_adaptive_threshold_kalman_filter, 0, 300, thr)
    g.update(adaptive_thr=np.float64(thr), adaptive_states=states_arr(st), adaptive_logdet=np.array(ldl),
             adaptive_P=P, adaptive_prev=np.float64(prev), adaptive_mtimes=np.array(mt))
    st, ldl, P, prev, mt = quiet(kf.run_adaptive_threshold_kalman_filter, 200, 330, thr + 1.5,
                                 P1.copy(), init_state)
    g.update(adaptive_warm_thr=np.float64(thr + 1.5), adaptive_warm_states=states_arr(st),
             adaptive_warm_logdet=np.array(ldl), adaptive_warm_P=P,
             adaptive_warm_prev=np.float64(prev), adaptive_warm_mtimes=np.array(mt))
    st, ldl, P, prev, mt = quiet(kf.run_no_update_kalman_filter, 0, 120)
    g.update(noupdate_states=states_arr(st), noupdate_logdet=np.array(ldl), noupdate_P=P,
             noupdate_prev=np.float64(prev), noupdate_mtimes=np.array(mt))

    for f in (20, 50, 120):
        kf.set_processing_frequency(f)
        st, ldl, P = quiet(kf.run_kalman_filter_scheduled, 0, 400, None, None, "greedy")
        g[f"sched_greedy_{f}_states"] = states_arr(st)
        g[f"sched_greedy_{f}_logdet"] = np.array(ldl)
        g[f"sched_greedy_{f}_P"] = P
    kf.set_processing_frequency(50)
    st, ldl, P = quiet(kf.run_kalman_filter_scheduled, 200, 400, P1.copy(), init_state, "greedy")
    g.update(sched_warm_states=states_arr(st), sched_warm_logdet=np.array(ldl), sched_warm_P=P)
    np.random.seed(7)
    st, ldl, P = quiet(kf.run_kalman_filter_scheduled, 0, 400, None, None, "random")
    g.update(sched_random_seed=np.int64(7), sched_random_states=states_arr(st),
             sched_random_logdet=np.array(ldl), sched_random_P=P)
    np.savez_compressed(os.path.join(OUT, "ca15_loops.npz"), **g)

    # ---------------- brute-force worker + driver (kf_workers.py:22-97, 1218-1392) ----------
    L = 10
    w0 = 150
    st, ld, P0, prev0 = quiet(kf.run_kalman_filter_full, 0, w0)
    init_state = st[-1]
    window = events[w0:w0 + L]
    xt = np.zeros(15)
    xt[0:6] = init_state[1:7]
    class_args = {name: getattr(kf, name) for name in (
        "get_state_transition_matrix", "get_process_noise_covariance_matrix", "predict_covariance",
        "get_gps_observation_matrix", "get_gps_measurement_noise_covariance_matrix",
        "get_imu_observation_matrix", "get_imu_measurement_noise_covariance_matrix",
        "calculate_kalman_gain")}
    t_end = events[w0 + L - 1][2]
    combos_idx = [c for k in range(1, L + 1) for c in itertools.combinations(range(L), k)]
    chunk = [tuple(window[i] for i in c) for c in combos_idx]
    res = quiet(kw.evaluate_combo_chunk_worker, chunk, xt, P0, class_args, init_state[0], t_end)
    assert len(res) == len(chunk)
    traces = np.full((len(res), L + 2), np.nan)
    xfin = np.zeros((len(res), 15))
    ntraj = np.zeros(len(res), dtype=np.int64)
    traj_last = np.zeros((len(res), 7))
    masks = np.zeros(len(res), dtype=np.uint64)
    for r, (metric, traj, combo, xf, _, log_det, num_used) in enumerate(res):
        traces[r, :len(log_det)] = log_det
        xfin[r] = xf
        ntraj[r] = len(traj)
        traj_last[r] = [float(v) for v in traj[-1]]
        masks[r] = sum(1 << i for i in combos_idx[r])
        assert metric == 0 and num_used == len(combo)
    w = {"ev_type": ty, "ev_time": tm, "ev_data": da, "w0": np.int64(w0), "L": np.int64(L),
         "init_state": np.array([float(v) for v in init_state]), "P0": P0, "t_end": np.float64(t_end),
         "masks": masks, "traces": traces, "x_final": xfin, "n_traj": ntraj, "traj_last": traj_last}

    # The driver itself on a smaller window.  Harness accommodation (the reference SOURCE is
    # untouched): the module-level name `Pool` it looks up is pointed at a thread-backed pool.
    # With real forked workers the driver's early `return` inside `with Pool(30)` hangs in this
    # container -- the forked workers inherit its SIGTERM handler (kf_workers.py:1235-1240) and
    # Pool.terminate() never finishes (the zombie problem its own comments describe).  The
    # worker function, enumeration, chunking and selection rule being pinned run unchanged.
    from multiprocessing.pool import ThreadPool
    kw.Pool = lambda processes=None: ThreadPool(2)
    Ld = 8
    maxes = np.nanmax(traces, axis=1)
    small = [r for r, c in enumerate(combos_idx) if max(c) < Ld]
    end_idx = w0 + Ld
    t_end_d = events[end_idx - 1][2]
    chunk_d = [tuple(window[i] for i in combos_idx[r]) for r in small]
    res_d = quiet(kw.evaluate_combo_chunk_worker, chunk_d, xt, P0, class_args, init_state[0], t_end_d)
    maxes_d = np.array([max(r[5]) for r in res_d])
    picks = []
    for q in (0.02, 0.3, 0.6):
        thr_d = float(np.quantile(maxes_d, q))
        out = quiet(kf.run_brute_force_kalman_filter_no_sampling_min_usage, w0, end_idx, thr_d,
                    P0.copy(), init_state)
        if out is None:
            picks.append((thr_d, -1, np.zeros(15), np.full(Ld + 2, np.nan)))
            continue
        sel = tuple(e[0] - w0 for e in out["selected_sensors"])
        lds = np.full(Ld + 2, np.nan)
        lds[:len(out["log_determinants"])] = out["log_determinants"]
        picks.append((thr_d, sum(1 << i for i in sel), out["final_state"], lds))
        assert out["num_measurements_used"] == len(sel)
    # and an infeasible threshold -> None
    out = quiet(kf.run_brute_force_kalman_filter_no_sampling_min_usage, w0, end_idx,
                float(maxes_d.min()) - 1.0, P0.copy(), init_state)
    assert out is None
    w.update(drv_L=np.int64(Ld), drv_thr=np.array([p[0] for p in picks]),
             drv_mask=np.array([p[1] for p in picks], dtype=np.int64),
             drv_final_state=np.array([p[2] for p in picks]),
             drv_logdet=np.array([p[3] for p in picks]),
             drv_infeasible_thr=np.float64(float(maxes_d.min()) - 1.0))
    np.savez_compressed(os.path.join(OUT, "ca15_bruteforce.npz"), **w)

    # ---------------- Scheduler (kf_workers.py:99-213) ----------------
    rng = np.random.default_rng(5)
    sch = kw.Scheduler()
    H_g, R_g = kf.get_gps_observation_matrix(), kf.get_gps_measurement_noise_covariance_matrix()
    H_i, R_i = kf.get_imu_observation_matrix(), kf.get_imu_measurement_noise_covariance_matrix()
    sig, post1, post3, posti1, postall, picks = [], [], [], [], [], []
    gps_ev, imu_ev = events[3], events[4]
    assert gps_ev[1] == "GPS" and imu_ev[1] == "IMU"
    queues = [[0, 1, 1], [1, 0, 1], [0], [1, 1], [0, 0, 1]]
    for _ in range(6):
        A = rng.normal(size=(15, 15))
        S = A @ A.T + 0.5 * np.eye(15)
        sig.append(S)
        post1.append(sch.cov_matrix([1], S, R_g, H_g, "cpu"))
        post3.append(sch.cov_matrix([1, 2, 3], S, R_g, H_g, "cpu"))
        posti1.append(sch.cov_matrix([1], S, R_i, H_i, "cpu"))
        postall.append(sch.cov_matrix(list(range(1, 16)), S, R_i, H_i, "cpu"))
        row = []
        for q in queues:
            qq = [gps_ev if c == 0 else imu_ev for c in q]
            # distinct tuples so .index() is unambiguous, as in the real queue
            qq = [(j, e[1], e[2] + j * 1e-3, e[3]) for j, e in enumerate(qq)]
            row.append(sch.greedy_schedule(qq, S, {"IMU": R_i, "GPS": R_g}, {"IMU": H_i, "GPS": H_g}, "cpu"))
        picks.append(row + [0] * 0)
    np.savez_compressed(os.path.join(OUT, "scheduler.npz"), Sigma=np.array(sig), post_gps1=np.array(post1),
                        post_gps3=np.array(post3), post_imu1=np.array(posti1), post_imu_all=np.array(postall),
                        queues=np.array([q + [-1] * (3 - len(q)) for q in queues]), greedy_pick=np.array(picks))

    # ---------------- hw5_2.py (8-state) ----------------
    events8 = make_indexed_sensor_data(320, seed=23)
    ty8, tm8, da8 = events_to_arrays(events8)
    events8 = arrays_to_events(ty8, tm8, da8)
    k8 = hw.KF_SensorFusion("gps.csv", "imu.csv")
    k8.indexed_sensor_data = events8
    s8 = quiet(k8.run_kalman_filter)
    d8 = quiet(k8.run_dead_reckoning_for_IMU)
    np.savez_compressed(os.path.join(OUT, "ca8_hw5_2.npz"), ev_type=ty8, ev_time=tm8, ev_data=da8,
                        kf_states=np.array(s8, dtype=np.float64), dr_states=np.array(d8, dtype=np.float64))
    notebook_bruteforce()
    accuracy_metrics()
    experiment_windows()
    csv_front_end()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
