"""Pins oracle/kf_oracle.py (the numpy restatement) to the golden vectors that
oracle/make_golden.py produced by running the UNMODIFIED reference.  CPU only."""
import itertools

import numpy as np
import pytest

from oracle import kf_oracle as ko
from oracle.make_golden import arrays_to_events, states_arr

TOL = dict(rtol=1e-11, atol=1e-11)


@pytest.fixture(scope="module")
def loops(golden):
    g = golden("ca15_loops.npz")
    g["events"] = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    return g


def test_full_cold(loops):
    m = ko.get_model("CA15")
    st, ld, P, prev = ko.run_full(m, loops["events"], 0, 400)
    np.testing.assert_allclose(states_arr(st), loops["full_cold_states"], **TOL)
    np.testing.assert_allclose(ld, loops["full_cold_logdet"], **TOL)
    np.testing.assert_allclose(P, loops["full_cold_P"], **TOL)
    assert prev == loops["full_cold_prev"]


def test_full_warm(loops):
    m = ko.get_model("CA15")
    st, ld, P, prev = ko.run_full(m, loops["events"], 200, 400, loops["full_warm_init_P"],
                                  tuple(loops["full_warm_init_state"]))
    np.testing.assert_allclose(states_arr(st), loops["full_warm_states"], **TOL)
    np.testing.assert_allclose(ld, loops["full_warm_logdet"], **TOL)
    np.testing.assert_allclose(P, loops["full_warm_P"], **TOL)
    assert prev == loops["full_warm_prev"]


def test_legacy(loops):
    m = ko.get_model("CA15")
    st, covs = ko.run_legacy(m, loops["events"], 0, 300)
    assert len(covs) == int(loops["legacy_n"])
    np.testing.assert_allclose(states_arr(st), loops["legacy_states"], **TOL)
    np.testing.assert_allclose(covs[-1], loops["legacy_P_last"], **TOL)
    np.testing.assert_allclose(covs[len(covs) // 2], loops["legacy_P_mid"], **TOL)


def test_adaptive_and_no_update(loops):
    m = ko.get_model("CA15")
    st, ld, P, prev, mt = ko.run_adaptive(m, loops["events"], 0, 300, float(loops["adaptive_thr"]))
    np.testing.assert_allclose(states_arr(st), loops["adaptive_states"], **TOL)
    np.testing.assert_allclose(ld, loops["adaptive_logdet"], **TOL)
    np.testing.assert_allclose(P, loops["adaptive_P"], **TOL)
    np.testing.assert_array_equal(np.array(mt), loops["adaptive_mtimes"])
    assert prev == loops["adaptive_prev"]
    st, ld, P, prev, mt = ko.run_adaptive(m, loops["events"], 200, 330, float(loops["adaptive_warm_thr"]),
                                          loops["full_warm_init_P"], tuple(loops["full_warm_init_state"]))
    np.testing.assert_allclose(states_arr(st), loops["adaptive_warm_states"], **TOL)
    np.testing.assert_allclose(ld, loops["adaptive_warm_logdet"], **TOL)
    np.testing.assert_array_equal(np.array(mt), loops["adaptive_warm_mtimes"])
    st, ld, P, prev, mt = ko.run_adaptive(m, loops["events"], 0, 120, None, never_update=True)
    np.testing.assert_allclose(states_arr(st), loops["noupdate_states"], **TOL)
    np.testing.assert_allclose(ld, loops["noupdate_logdet"], **TOL)
    np.testing.assert_allclose(P, loops["noupdate_P"], **TOL)
    np.testing.assert_array_equal(np.array(mt), loops["noupdate_mtimes"])


@pytest.mark.parametrize("f", [20, 50, 120])
def test_scheduled_greedy(loops, f):
    m = ko.get_model("CA15")
    st, ld, P = ko.run_scheduled(m, loops["events"], f, 0, 400)
    np.testing.assert_allclose(states_arr(st), loops[f"sched_greedy_{f}_states"], **TOL)
    np.testing.assert_allclose(ld, loops[f"sched_greedy_{f}_logdet"], **TOL)
    np.testing.assert_allclose(P, loops[f"sched_greedy_{f}_P"], **TOL)


def test_scheduled_warm_and_random(loops):
    m = ko.get_model("CA15")
    st, ld, P = ko.run_scheduled(m, loops["events"], 50, 200, 400, loops["full_warm_init_P"],
                                 tuple(loops["full_warm_init_state"]))
    np.testing.assert_allclose(states_arr(st), loops["sched_warm_states"], **TOL)
    np.testing.assert_allclose(ld, loops["sched_warm_logdet"], **TOL)
    np.random.seed(int(loops["sched_random_seed"]))
    st, ld, P = ko.run_scheduled(m, loops["events"], 50, 0, 400, selection="random",
                                 rng_choice=np.random.choice)
    np.testing.assert_allclose(states_arr(st), loops["sched_random_states"], **TOL)
    np.testing.assert_allclose(P, loops["sched_random_P"], **TOL)


@pytest.fixture(scope="module")
def brute(golden):
    g = golden("ca15_bruteforce.npz")
    g["events"] = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    return g


def test_worker_all_subsets(brute):
    m = ko.get_model("CA15")
    L, w0 = int(brute["L"]), int(brute["w0"])
    window = brute["events"][w0:w0 + L]
    init = brute["init_state"]
    xt = np.zeros(15)
    xt[0:6] = init[1:7]
    order = list(ko.enumerate_subsets(L))
    assert [ko.subset_mask(c) for c in order] == [int(v) for v in brute["masks"]]
    types = [0 if e[1] == "GPS" else 1 for e in window]
    times = [e[2] for e in window]
    for r in range(0, len(order), 7):  # every 7th subset keeps the CPU suite short
        combo = tuple(window[i] for i in order[r])
        traj, x, P, lds = ko.worker_combo(m, combo, xt, brute["P0"], init[0], float(brute["t_end"]))
        ref = brute["traces"][r]
        ref = ref[~np.isnan(ref)]
        np.testing.assert_allclose(lds, ref, **TOL)
        np.testing.assert_allclose(x, brute["x_final"][r], **TOL)
        assert len(traj) == int(brute["n_traj"][r])
        np.testing.assert_allclose([float(v) for v in traj[-1]], brute["traj_last"][r], **TOL)
        # the covariance-only trace is the same list (measurement independence)
        cov_only = ko.schedule_trace(m, types, times, int(brute["masks"][r]), brute["P0"], init[0],
                                     float(brute["t_end"]))
        np.testing.assert_allclose(cov_only, ref, **TOL)


def test_driver_selection(brute):
    m = ko.get_model("CA15")
    Ld, w0 = int(brute["drv_L"]), int(brute["w0"])
    window = brute["events"][w0:w0 + Ld]
    init = brute["init_state"]
    xt = np.zeros(15)
    xt[0:6] = init[1:7]
    t_end = window[-1][2]
    for thr, mask, xf, lds in zip(brute["drv_thr"], brute["drv_mask"], brute["drv_final_state"],
                                  brute["drv_logdet"]):
        out = ko.brute_force_select(m, window, xt, brute["P0"], init[0], t_end, float(thr))
        if int(mask) < 0:  # reference returned None: the max includes the shared initial term
            assert out is None
            continue
        assert out is not None and ko.subset_mask(out["selected_indices"]) == int(mask)
        np.testing.assert_allclose(out["final_state"], xf, **TOL)
        np.testing.assert_allclose(out["log_determinants"], lds[~np.isnan(lds)], **TOL)
    assert ko.brute_force_select(m, window, xt, brute["P0"], init[0], t_end,
                                 float(brute["drv_infeasible_thr"])) is None


def test_lex_rank_key_matches_enumeration():
    L = 7
    order = [ko.subset_mask(c) for c in ko.enumerate_subsets(L)]
    assert sorted(order, key=lambda mk: ko.lex_rank_key(mk, L)) == order


def test_scheduler(golden):
    g = golden("scheduler.npz")
    m = ko.get_model("CA15")
    for k, S in enumerate(g["Sigma"]):
        np.testing.assert_allclose(ko.cov_matrix([1], S, m.R_gps(), m.H_gps()), g["post_gps1"][k], **TOL)
        np.testing.assert_allclose(ko.cov_matrix([1, 2, 3], S, m.R_gps(), m.H_gps()), g["post_gps3"][k], **TOL)
        np.testing.assert_allclose(ko.cov_matrix([1], S, m.R_imu(), m.H_imu()), g["post_imu1"][k], **TOL)
        np.testing.assert_allclose(ko.cov_matrix(list(range(1, 16)), S, m.R_imu(), m.H_imu()),
                                   g["post_imu_all"][k], **TOL)
        for q, pick in zip(g["queues"], g["greedy_pick"][k]):
            queue = [(j, "GPS" if c == 0 else "IMU", float(j), None) for j, c in enumerate(q) if c >= 0]
            assert ko.greedy_schedule(m, queue, S) == int(pick)


def test_hw5_2(golden):
    g = golden("ca8_hw5_2.npz")
    ev = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    m = ko.get_model("CA8")
    np.testing.assert_allclose(np.array(ko.hw5_2_run(m, ev), dtype=float), g["kf_states"], **TOL)
    np.testing.assert_allclose(np.array(ko.hw5_2_dead_reckoning(m, ev), dtype=float), g["dr_states"], **TOL)


def test_batch_matches_event_loop(loops):
    """The packed run_batch (what the C ABI computes) equals the event loop on the same stream."""
    m = ko.get_model("CA15")
    ev = loops["events"][3:120]  # starts at the first GPS fix, no negative dt inside
    assert ev[0][1] == "GPS"
    st, ld, P, _ = ko.run_full(m, ev, 0, len(ev))
    times = np.array([e[2] for e in ev])
    dts = np.diff(times, prepend=times[0])
    types = np.array([0 if e[1] == "GPS" else 1 for e in ev], dtype=np.uint8)
    payload = np.stack([ko.event_payload(m, e[1], e[3]) for e in ev])[:, :, None]
    x0 = np.zeros((1, 15))
    x0[0, :3] = [ev[0][3]["easting"], ev[0][3]["northing"], ev[0][3]["altitude"]]
    out = ko.run_batch(m, types, dts, payload, x0, m.P0()[None])
    np.testing.assert_allclose(out["traj"][:, :6, 0], states_arr(st)[1:, 1:], **TOL)
    np.testing.assert_allclose(out["logdet"][:, 0], ld[1:], **TOL)
    np.testing.assert_allclose(out["P"][0], P, **TOL)


def test_notebook_product_search(golden):
    """The product-of-queues / RMSE brute force that only exists in the reference's notebook."""
    g = golden("nb_bruteforce.npz")
    ev = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    m = ko.get_model("CA15")
    gps = [e for e in ev if e[1] == "GPS"]
    tt = np.array([e[2] for e in gps])
    tp = np.array([[e[3]["easting"], e[3]["northing"], e[3]["altitude"]] for e in gps])
    for tag in "ab":
        s0, s1 = (int(v) for v in g[f"{tag}_range"])
        queues = ko.bucket_queues(ev, s0, s1, float(g[f"{tag}_freq"]))
        out = ko.product_search(m, queues, tt, tp)
        assert [e[0] for e in out["selected_sensors"]] == [int(v) for v in g[f"{tag}_selected"]]
        np.testing.assert_allclose(out["accuracy_metric"], g[f"{tag}_metric"], **TOL)
        np.testing.assert_allclose(out["final_state"], g[f"{tag}_final_state"], **TOL)
        np.testing.assert_allclose(out["final_covariance"], g[f"{tag}_final_cov"], **TOL)
        np.testing.assert_allclose(states_arr(out["trajectory"]), g[f"{tag}_trajectory"], **TOL)


def test_lerp_is_bit_identical_to_scipy():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(5)
    xp = np.sort(rng.uniform(1.6977e9, 1.6977e9 + 50, 40))
    fp = rng.normal(4.6e6, 30.0, 40)
    x = np.concatenate([rng.uniform(xp[0] - 5, xp[-1] + 5, 200), xp[::3]])  # incl. knots and extrapolation
    want = interp1d(xp, fp, kind="linear", fill_value="extrapolate")(x)
    np.testing.assert_array_equal(ko.lerp_extrapolate(x, xp, fp), want)


def test_accuracy_metrics(golden):
    """calculate_accuracy_metrics (kf_workers.py:1162-1216) incl. the < 2-sample fallback."""
    g = golden("ca15_accuracy.npz")
    ev = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    m = ko.get_model("CA15")
    gt, _, _, _ = ko.run_full(m, ev, 0, 400)
    for tag, key in (("cand_states", "rmse"), ("tiny_states", "tiny_rmse")):
        cand = [tuple(r) for r in g[tag]]
        rmse, interp = ko.accuracy_metrics(cand, gt)
        np.testing.assert_allclose(rmse, g[key], rtol=1e-9)
        np.testing.assert_allclose(interp, g["ground_truth_interp" if key == "rmse" else "tiny_interp"], rtol=1e-11)


# ---- full-LENGTH golden vectors (oracle/make_golden_long.py): the oracle itself is held to the reference
# over 120 000 events of the full filter and over the whole 2 808 s stream of the 10 Hz scheduled filter
def _long_sample(a, stride):
    a = np.asarray(a)
    return np.concatenate([a[:-8:stride], a[-8:]], axis=0)


def test_oracle_full_filter_at_120k_events(golden):
    from sensorfusion_kalmanfilter_b200 import synth
    g = golden("ca15_full_120k.npz")
    n, stride = int(g["n_events"]), int(g["stride"])
    ev = synth.make_indexed_sensor_data(n + 3, seed=int(g["base_seed"]))[3:]
    st, ld, P, prev = ko.run_full(ko.get_model("CA15"), ev, 0, len(ev))
    assert len(st) == int(g["n_states"]) and prev == float(g["prev_time"])
    np.testing.assert_allclose(_long_sample(states_arr(st), stride), g["states"], rtol=1e-10, atol=1e-10)
    ld = np.array(ld)
    np.testing.assert_allclose(_long_sample(ld, stride), g["logdet"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose([ld.max(), ld.sum(), ld[-1]], g["agg"], rtol=1e-10)
    np.testing.assert_allclose(P, g["P"], rtol=1e-10, atol=1e-12)


def test_oracle_scheduled_filter_at_config2_length(golden):
    from sensorfusion_kalmanfilter_b200 import streams, synth
    g = golden("ca15_config2_full_length.npz")
    n, stride, seed = int(g["n_events"]), int(g["stride"]), int(g["seeds"][1])
    ev = synth.make_indexed_sensor_data(n + 3, seed=int(g["base_seed"]))[3:]
    ps = streams.PackedStream.from_events(ev)
    ev = arrays_to_events(ps.types, ps.times, ps.data + synth.seed_noise(ps.types, seed))
    st, ld, P = ko.run_scheduled(ko.get_model("CA15"), ev, 10, 0, len(ev))
    assert len(st) == int(g[f"r10_s{seed}_n_states"])
    sel = np.searchsorted(ps.times, states_arr(st)[1:, 0])
    np.testing.assert_array_equal(sel, np.cumsum(g["r10_sel_delta"].astype(np.int64)))
    np.testing.assert_allclose(_long_sample(states_arr(st), stride), g[f"r10_s{seed}_states"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(P, g[f"r10_s{seed}_P"], rtol=1e-10, atol=1e-12)
