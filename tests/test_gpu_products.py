"""Product-of-queues brute force with the RMSE objective (KF_SensorFusion.ipynb:804-954) and the
accuracy metric (kf_workers.py:1162-1216) on the GPU, through the C ABI, against the golden vectors
of the unmodified reference (notebook cells exec'd by oracle/make_golden.py) and the oracle."""
import numpy as np
import pytest

from oracle import kf_oracle as ko
from oracle.make_golden import arrays_to_events, states_arr

pytestmark = pytest.mark.gpu

R = 1e-9


def close(a, b, tol=R):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b))), np.max(np.abs(a - b))


@pytest.fixture(scope="module")
def nb(golden):
    g = golden("nb_bruteforce.npz")
    ev = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    return g, ev


def test_notebook_brute_force_mirror(nb):
    """The notebook's own call: k.run_brute_force_kalman_filter(start, end)."""
    from sensorfusion_kalmanfilter_b200.notebook import KF_SensorFusion
    g, ev = nb
    k = KF_SensorFusion("gps.csv", "imu.csv")
    k.indexed_sensor_data = ev
    k.utm_data = [e[3] for e in ev if e[1] == "GPS"]
    for tag in "ab":
        k.set_processing_frequency(float(g[f"{tag}_freq"]))
        s0, s1 = (int(v) for v in g[f"{tag}_range"])
        out = k.run_brute_force_kalman_filter(s0, s1)
        assert [e[0] for e in out["selected_sensors"]] == [int(v) for v in g[f"{tag}_selected"]]  # bit-exact
        close(out["accuracy_metric"], g[f"{tag}_metric"])
        close(out["final_state"], g[f"{tag}_final_state"])
        close(out["final_covariance"], g[f"{tag}_final_cov"])
        tr = states_arr(out["trajectory"])
        np.testing.assert_array_equal(tr[:, 0], g[f"{tag}_trajectory"][:, 0])
        close(tr[:, 1:], g[f"{tag}_trajectory"][:, 1:])
        assert out["runner_up_metric"] > out["accuracy_metric"]
        close(k.calculate_accuracy_metrics(out["trajectory"]), g[f"{tag}_rmse_of_best"])
    # overwrite_sampling_freq takes precedence over processing_frequency (:838-841)
    k.set_processing_frequency(5.0)
    out = k.run_brute_force_kalman_filter(0, 40, overwrite_sampling_freq=50)
    assert [e[0] for e in out["selected_sensors"]] == [int(v) for v in g["a_selected"]]
    # no GPS in range -> no queue -> nothing selected (:932-933 never fires)
    imu_only = [e for e in ev if e[1] == "IMU"][:10]
    k2 = KF_SensorFusion("gps.csv", "imu.csv")
    k2.indexed_sensor_data = imu_only
    k2.utm_data = k.utm_data
    k2.set_processing_frequency(50)
    out = k2.run_brute_force_kalman_filter(0, 10)
    assert out["selected_sensors"] is None and out["accuracy_metric"] == float("inf")


@pytest.mark.parametrize("mode", ["auto", "dense", "decoupled", "symmetric"])
def test_score_products_every_sequence_vs_oracle(nb, mode):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    from sensorfusion_kalmanfilter_b200.streams import PackedStream
    g, ev = nb
    m = ko.get_model("CA15")
    gps = [e for e in ev if e[1] == "GPS"]
    tt = np.array([e[2] for e in gps])
    tp = np.array([[e[3]["easting"], e[3]["northing"], e[3]["altitude"]] for e in gps])
    queues = ko.bucket_queues(ev, 5, 44, 70.0)
    want = ko.product_search(m, queues, tt, tp)
    eng = Engine("CA15")
    flat = [e for q in queues for e in q]
    ps = PackedStream.from_events(flat)
    ref = eng.track_lerp(tt, tp, ps.times)
    ref_o = np.stack([ko.lerp_extrapolate(ps.times, tt, tp[:, c]) for c in range(3)], axis=1)
    np.testing.assert_array_equal(ref.cpu().numpy(), ref_o)  # bit-identical interpolation
    qlen = [len(q) for q in queues]
    best, per = eng.score_products(qlen, ps.types, ps.times, ps.payload("CA15"), ref, np.zeros(15), ko.nb_P0(),
                                   want_per_combo=True, mode=mode)
    per = per.cpu().numpy()
    assert per.shape == want["metrics"].shape
    close(per, want["metrics"])
    assert best["index"] == want["combo_index"]
    close(best["score"], want["accuracy_metric"])
    close(best["runner_up_score"], np.sort(want["metrics"])[1])
    assert best["n_scored"] == len(per)
    # sharded ranges (one per GPU in the multi-GPU search): the reduce of the shard winners is the winner
    total = len(per)
    cuts = [0, total // 3, total // 3 + 1, total]
    shard = [eng.score_products(qlen, ps.types, ps.times, ps.payload("CA15"), ref, np.zeros(15), ko.nb_P0(),
                                first=a, count=b - a, mode=mode)[0] for a, b in zip(cuts[:-1], cuts[1:])]
    win = min(shard, key=lambda b: (b["score"], b["index"]))
    assert win["index"] == want["combo_index"]


def test_score_products_other_models_and_warm_start():
    """CV6 / CA8 sequences from a non-trivial (x0, dense P0) against the oracle step by step, incl.
    one and two windows (no shared prefix to walk)."""
    import itertools
    from sensorfusion_kalmanfilter_b200.engine import Engine
    rng = np.random.default_rng(7)
    for name, qlen in itertools.product(("CV6", "CA8"), ([3, 1, 4, 2, 3], [5], [2, 3])):
        m = ko.get_model(name)
        E = sum(qlen)
        times = 100.0 + np.sort(rng.uniform(0, 0.5, E))
        types = rng.integers(0, 2, E).astype(np.uint8)
        pay = rng.normal(0, 1.0, (E, m.d))
        ref = rng.normal(0, 1.0, (E, 3))
        x0 = rng.normal(0, 1, m.n)
        A = rng.normal(0, 1, (m.n, m.n))
        P0 = A @ A.T + 5 * np.eye(m.n)
        eng = Engine(name)
        best, per = eng.score_products(qlen, types, times, pay, ref, x0, P0, want_per_combo=True)
        per = per.cpu().numpy()
        off = np.concatenate(([0], np.cumsum(qlen)[:-1]))
        want = []
        for digits in itertools.product(*[range(q) for q in qlen]):
            x, P, prev, sse = x0.copy(), P0.copy(), None, 0.0
            for w, dgt in enumerate(digits):
                e = off[w] + dgt
                dt = times[e] - prev if prev is not None else 0.0
                x, P = ko.step(m, x, P, int(types[e]), dt, pay[e])
                sse += np.linalg.norm(x[:3] - ref[e]) ** 2
                prev = times[e]
            want.append(np.sqrt(sse / len(qlen)))
        want = np.array(want)
        close(per, want)
        assert best["index"] == int(np.argmin(want))
        # a ragged sub-range that starts and ends inside a thread's block of sequences
        if len(want) > 6:
            b2, p2 = eng.score_products(qlen, types, times, pay, ref, x0, P0, first=2, count=len(want) - 5,
                                        want_per_combo=True)
            close(p2.cpu().numpy(), want[2:len(want) - 3])
            assert b2["index"] == 2 + int(np.argmin(want[2:len(want) - 3]))


def test_score_products_argument_errors():
    from sensorfusion_kalmanfilter_b200 import _lib
    from sensorfusion_kalmanfilter_b200.engine import Engine
    eng = Engine("CA15")
    z = np.zeros
    with pytest.raises(_lib.B200KFError):  # an empty queue: no sequence exists
        eng.score_products([2, 0], z(2, np.uint8), z(2), z((2, 9)), z((2, 3)), z(15), np.eye(15))
    with pytest.raises(_lib.B200KFError):  # range beyond the product of the queue lengths
        eng.score_products([2, 2], z(4, np.uint8), z(4), z((4, 9)), z((4, 3)), z(15), np.eye(15), first=3, count=2)


def test_accuracy_metrics_mirror(golden):
    """kf_workers.calculate_accuracy_metrics on the device vs the unmodified reference."""
    from sensorfusion_kalmanfilter_b200.kf_workers import KF_SensorFusion
    g = golden("ca15_accuracy.npz")
    k = KF_SensorFusion("gps.csv", "imu.csv")
    k.indexed_sensor_data = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    assert k.calculate_accuracy_metrics([(0.0, 0, 0, 0)]) is None  # no ground truth yet (:1170-1172)
    k.run_kalman_filter_full(0, 400)
    assert k.calculate_accuracy_metrics([]) is None
    k.set_processing_frequency(30)
    st, _, _ = k.run_kalman_filter_scheduled(40, 360, None, None, "greedy")
    m = k.calculate_accuracy_metrics(st)
    close(m["total_position_rmse"], g["rmse"])
    close(m["ground_truth_interp"], g["ground_truth_interp"])
    close(m["euclidean_errors"], g["euclidean_errors"], 1e-8)
    m2 = k.calculate_accuracy_metrics([tuple(r) for r in g["tiny_states"]])
    close(m2["total_position_rmse"], g["tiny_rmse"])
    close(m2["ground_truth_interp"], g["tiny_interp"])


def test_rmse_vs_track_batched_many_filters_and_long_single():
    """The time-segmented reduction: many filters x few steps and one filter x many steps."""
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine
    eng = Engine("CV6")
    rng = np.random.default_rng(3)
    for T, N in ((7, 5000), (20000, 1), (300, 33)):
        traj = rng.normal(0, 2, (T, 6, N))
        ref = rng.normal(0, 2, (T, 3))
        got = eng.rmse_vs_track(torch.as_tensor(traj), ref).cpu().numpy()
        e = traj[:, :3, :] - ref[:, :, None]
        want = np.sqrt(np.mean(np.sum(e * e, axis=1), axis=0))
        np.testing.assert_allclose(got, want, rtol=1e-12)
