"""GPU parity of b200kf_score_schedules / b200kf_cov_update against the golden vectors produced by
the unmodified reference (tests/golden/ca15_bruteforce.npz, scheduler.npz) and against the oracle."""
import numpy as np
import pytest

from oracle import kf_oracle as ko
from oracle.make_golden import arrays_to_events

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def brute(golden):
    g = golden("ca15_bruteforce.npz")
    g["events"] = arrays_to_events(g["ev_type"], g["ev_time"], g["ev_data"])
    return g


def window_arrays(g, L):
    w0 = int(g["w0"])
    ev = g["events"][w0:w0 + L]
    types = np.array([0 if e[1] == "GPS" else 1 for e in ev], dtype=np.uint8)
    times = np.array([e[2] for e in ev], dtype=np.float64)
    return types, times


@pytest.mark.parametrize("mode", ["auto", "dense", "decoupled", "symmetric"])
def test_all_subsets_match_reference_traces(brute, mode):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    L = int(brute["L"])
    types, times = window_arrays(brute, L)
    eng = Engine("CA15")
    best, per = eng.score_schedules(types, times, brute["init_state"][0], float(brute["t_end"]), brute["P0"],
                                    first=0, count=1 << L, threshold=0.0, want_per_schedule=True, mode=mode)
    per = per.cpu().numpy()
    tr = brute["traces"]
    masks = brute["masks"].astype(np.int64)
    ref_max = np.nanmax(tr, axis=1)
    ref_sum = np.nansum(tr, axis=1)
    ref_cnt = np.sum(~np.isnan(tr), axis=1)
    ref_last = np.array([row[~np.isnan(row)][-1] for row in tr])
    got = per[masks]
    assert np.max(np.abs(got[:, 0] - ref_max) / np.maximum(1, np.abs(ref_max))) < 1e-9
    assert np.max(np.abs(got[:, 1] - ref_sum) / np.maximum(1, np.abs(ref_sum))) < 1e-9
    assert np.max(np.abs(got[:, 2] - ref_last) / np.maximum(1, np.abs(ref_last))) < 1e-9
    np.testing.assert_array_equal(got[:, 3], ref_cnt)
    # argmin of the sum, ties -> lowest index (bit-exact index)
    order = np.lexsort((masks, ref_sum))
    assert best["argmin_sum_index"] == int(masks[order[0]])


def test_driver_selection_bit_exact(brute):
    """First feasible schedule in the reference's (cardinality, lexicographic) order."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    Ld = int(brute["drv_L"])
    types, times = window_arrays(brute, Ld)
    eng = Engine("CA15")
    sgn, lad = np.linalg.slogdet(brute["P0"])  # term 0 exactly as the caller (and the reference) has it
    for thr, mask in zip(brute["drv_thr"], brute["drv_mask"]):
        best, _ = eng.score_schedules(types, times, brute["init_state"][0], float(times[-1]), brute["P0"],
                                      threshold=float(thr), initial_logdet=float(sgn * lad))
        if int(mask) < 0:
            assert best["first_feasible_index"] is None and best["n_feasible"] == 0
        else:
            assert best["first_feasible_index"] == int(mask)
    best, _ = eng.score_schedules(types, times, brute["init_state"][0], float(times[-1]), brute["P0"],
                                  threshold=float(brute["drv_infeasible_thr"]))
    assert best["first_feasible_index"] is None


def test_sharded_ranges_reduce_to_the_same_answer(brute):
    """Index-range sharding (what each GPU rank does) + the host reduce equals one full search."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    from sensorfusion_kalmanfilter_b200.distributed import merge_best, shard_range
    L = int(brute["L"])
    types, times = window_arrays(brute, L)
    eng = Engine("CA15")
    maxes = np.nanmax(brute["traces"], axis=1)
    thr = float(np.quantile(maxes, 0.4))
    args = (types, times, brute["init_state"][0], float(brute["t_end"]), brute["P0"])
    full, _ = eng.score_schedules(*args, threshold=thr)
    parts = []
    for r in range(5):
        first, count = shard_range(1 << L, 5, r)
        b, _ = eng.score_schedules(*args, first=first, count=count, threshold=thr)
        parts.append(b)
    merged = merge_best(parts)
    for k in ("argmin_sum_index", "first_feasible_index", "n_feasible"):
        assert merged[k] == full[k]
    assert merged["argmin_sum_score"] == full["argmin_sum_score"]


@pytest.mark.parametrize("name", ["CV6", "BIAS9", "CA8", "CVBU6"])
def test_other_models_match_oracle(name):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    m = ko.get_model(name)
    rng = np.random.default_rng(3)
    L = 9
    types = (rng.random(L) < 0.7).astype(np.uint8)
    times = 100.0 + np.cumsum(0.004 + 0.004 * rng.random(L))
    A = rng.normal(size=(m.n, m.n))
    P0 = A @ A.T + np.eye(m.n) if name != "CA8" else m.P0()
    eng = Engine(name)
    best, per = eng.score_schedules(types, times, 99.99, float(times[-1]) + 0.01, P0, want_per_schedule=True)
    per = per.cpu().numpy()
    sums = []
    for s in range(1 << L):
        tr = ko.schedule_trace(m, types, times, s, P0, 99.99, float(times[-1]) + 0.01)
        assert abs(per[s, 0] - max(tr)) < 1e-9 * max(1, abs(max(tr)))
        assert abs(per[s, 1] - sum(tr)) < 1e-9 * max(1, abs(sum(tr)))
        assert per[s, 3] == len(tr)
        sums.append(sum(tr))
    assert best["argmin_sum_index"] == 1 + int(np.argmin(sums[1:]))


def test_cov_update_matches_reference_scheduler(golden):
    from sensorfusion_kalmanfilter_b200.engine import cov_update
    g = golden("scheduler.npz")
    m = ko.get_model("CA15")
    S = g["Sigma"]
    for rows, H, R, key in (([0], m.H_gps(), m.R_gps(), "post_gps1"), ([0, 1, 2], m.H_gps(), m.R_gps(), "post_gps3"),
                            ([0], m.H_imu(), m.R_imu(), "post_imu1"),
                            (list(range(15)), m.H_imu(), m.R_imu(), "post_imu_all")):
        post, tr = cov_update(S, rows, H, R)
        ref = g[key]
        assert np.max(np.abs(post.cpu().numpy() - ref)) / np.max(np.abs(ref)) < 1e-9
        np.testing.assert_allclose(tr.cpu().numpy(), np.trace(ref, axis1=1, axis2=2), rtol=1e-9)


@pytest.mark.parametrize("name,mode", [("CA15", "auto"), ("CA15", "decoupled"), ("CV6", "auto"), ("CA8", "auto")])
def test_prefix_sharing_kernel_is_bit_identical_to_flat_kernel(name, mode, monkeypatch):
    """Aligned blocks of >= 2^14 schedules take the prefix-sharing (DFS) kernel; B200KF_TREE=0 keeps
    one thread per schedule.  Same arithmetic in the same order -> every per-schedule aggregate, the
    argmin, the first feasible schedule and the counts agree exactly, for the full range, a ragged
    range (flat head + tree blocks + flat tail) and the shards of a 2-GPU / 8-GPU split; a sample of
    schedules is also held to the oracle's trace."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    m = ko.get_model(name)
    L = 17
    rng = np.random.default_rng(23)
    types = np.where(np.arange(L) % 5 == 2, 0, 1).astype(np.uint8)
    times = 500.0 + np.cumsum(0.004 + 0.002 * rng.random(L))
    times[9] = times[7] - 1e-4  # an out-of-order event: skipped when it would step back (kf_workers.py:38-40)
    P0 = m.P0() * 1e-2
    t_prev, t_end = 499.99, float(times[-1]) + 0.01
    monkeypatch.setenv("B200KF_TREE", "0")
    flat_eng = Engine(name)
    monkeypatch.setenv("B200KF_TREE", "1")
    tree_eng = Engine(name)
    bf, pf = flat_eng.score_schedules(types, times, t_prev, t_end, P0, want_per_schedule=True, mode=mode)
    thr = float(np.median(pf[:, 0].cpu().numpy()[1:]))
    for first, count in ((0, 1 << L), (12345, 100001), (1 << 16, 1 << 16), (3 << 14, 1 << 14), (5, 300)):
        bf, pf = flat_eng.score_schedules(types, times, t_prev, t_end, P0, first=first, count=count, threshold=thr,
                                          want_per_schedule=True, mode=mode)
        bt, pt = tree_eng.score_schedules(types, times, t_prev, t_end, P0, first=first, count=count, threshold=thr,
                                          want_per_schedule=True, mode=mode)
        np.testing.assert_array_equal(pt.cpu().numpy(), pf.cpu().numpy())
        assert bt == bf
    per = pt.cpu().numpy()  # last range: [5, 305)
    for s in (5, 77, 304):
        tr = ko.schedule_trace(m, types, times, s, P0, t_prev, t_end)
        row = per[s - 5]
        assert abs(row[0] - np.max(tr)) <= 1e-9 * max(1, abs(np.max(tr)))
        assert abs(row[1] - np.sum(tr)) <= 1e-9 * max(1, abs(np.sum(tr)))
        assert row[3] == len(tr)
    bt, pt = tree_eng.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, want_per_schedule=True, mode=mode)
    per = pt.cpu().numpy()
    for s in (1, 0x1ABCD, (1 << L) - 1, 1 << 16):
        tr = ko.schedule_trace(m, types, times, s, P0, t_prev, t_end)
        assert abs(per[s][1] - np.sum(tr)) <= 1e-9 * max(1, abs(np.sum(tr)))
        assert abs(per[s][2] - tr[-1]) <= 1e-9 * max(1, abs(tr[-1]))


@pytest.mark.parametrize("L,rows", [(22, True), (25, False)])
def test_deep_trees_match_the_flat_kernel(L, rows, monkeypatch):
    """Deeper depth-first walks than the 17- and 20-event cases reach: 5 levels at 22 events (the two deepest in shared
    memory, three in local memory), the maximum of 8 at 25.  Same per-schedule rows (22 events) and the same winner,
    first feasible schedule and counts (both) as one thread per schedule."""
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine
    m = ko.get_model("CA15")
    rng = np.random.default_rng(41)
    types = np.where(np.arange(L) % 6 == 4, 0, 1).astype(np.uint8)
    times = 1000.0 + np.cumsum(0.005 + 0.0005 * rng.random(L))
    P0 = m.P0() * 1e-8
    t_prev, t_end = 999.995, float(times[-1])
    monkeypatch.setenv("B200KF_TREE", "0")
    flat = Engine("CA15")
    monkeypatch.setenv("B200KF_TREE", "1")
    tree = Engine("CA15")
    _, probe = tree.score_schedules(types, times, t_prev, t_end, P0, first=1, count=(1 << 16) - 1, want_per_schedule=True)
    thr = float(probe[:, 0].median().item())
    bt, pt = tree.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, want_per_schedule=rows)
    bf, pf = flat.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, want_per_schedule=rows)
    assert bt == bf and 0 < bt["n_feasible"] < (1 << L) - 1
    if rows:
        assert torch.equal(pt, pf)


def test_bf_schedule_tree_matches_restatement():
    """a17: the exhaustive multi-day schedule tree of sensor_scheduling_algorithm.ipynb:213-291 with the
    sizes of its first cell (n = 2 states, G = 5 sensors, K = 3 per day, horizon 3) and a larger case.
    The sketch itself is not runnable, so parity is against the oracle's loop-for-loop restatement."""
    from sensorfusion_kalmanfilter_b200.kf_workers import Scheduler
    rng = np.random.default_rng(11)
    for n, G, k, days in ((2, 5, 3, 3), (6, 6, 2, 3), (15, 4, 1, 4)):
        H = rng.random((G, n))
        A = rng.normal(size=(n, n))
        P = A @ A.T + np.eye(n)
        z = 1.0 + rng.random((days, G)) * 10.0
        want_sig, want_S = ko.bf_schedule(P, H, z, days, k)
        got_sig, got_S, times = Scheduler().BF_schedule(z, days, P=P, H=H, total_gauges=G, selected_gauges=k)
        assert [tuple(s) for s in got_S] == [tuple(s) for s in want_S]  # the same path, bit-exact
        assert len(times) == days and all(b >= a for a, b in zip(times, times[1:]))
        for a, b in zip(got_sig, want_sig):
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b))


def test_config4_full_size_properties(monkeypatch):
    """BASELINE configs[3] at its full size: all 2^20 subsets of a 20-event window (CA15).  The oracle
    cannot replay a million schedules, so: (a) the prefix-sharing kernel and the flat kernel agree on
    the winner, the first feasible schedule and the counts; (b) the 8 contiguous shards of an 8-GPU run
    reduce to the same answer (distributed.merge_best is what every rank applies after the all-gather);
    (c) the winners themselves are replayed through the oracle; (d) feasibility is monotone in the
    threshold; (e) a schedule's sum of log-dets equals that of the same events walked by the filter loop."""
    from sensorfusion_kalmanfilter_b200.distributed import merge_best, shard_range
    from sensorfusion_kalmanfilter_b200.engine import Engine
    m = ko.get_model("CA15")
    L = 20
    rng = np.random.default_rng(3)
    types = np.where(np.arange(L) % 7 == 3, 0, 1).astype(np.uint8)
    times = 1000.0 + np.cumsum(0.005 + 0.0005 * rng.random(L))
    P0 = m.P0() * 1e-8  # small enough that the covariance GROWS between measurements: the max is not term 0
    t_prev, t_end = 999.995, float(times[-1])
    monkeypatch.setenv("B200KF_TREE", "0")
    flat = Engine("CA15")
    monkeypatch.setenv("B200KF_TREE", "1")
    tree = Engine("CA15")
    probe, _ = tree.score_schedules(types, times, t_prev, t_end, P0, first=1, count=4096, want_per_schedule=True)
    thr = float(np.quantile(_.cpu().numpy()[:, 0], 0.4))
    bt, _ = tree.score_schedules(types, times, t_prev, t_end, P0, threshold=thr)
    bf, _ = flat.score_schedules(types, times, t_prev, t_end, P0, threshold=thr)
    assert bt == bf and bt["n_feasible"] > 0
    parts = []
    for r in range(8):
        first, count = shard_range(1 << L, 8, r)
        parts.append(tree.score_schedules(types, times, t_prev, t_end, P0, first=first, count=count, threshold=thr)[0])
    merged = merge_best(parts)
    for k in ("argmin_sum_index", "first_feasible_index", "n_feasible", "n_near_threshold"):
        assert merged[k] == bt[k], k
    for s in (bt["argmin_sum_index"], bt["first_feasible_index"]):
        tr = ko.schedule_trace(m, types, times, s, P0, t_prev, t_end)
        one, per = tree.score_schedules(types, times, t_prev, t_end, P0, first=s, count=1, threshold=thr,
                                        want_per_schedule=True)
        row = per.cpu().numpy()[0]
        assert abs(row[1] - np.sum(tr)) <= 1e-9 * max(1.0, abs(np.sum(tr)))
        assert abs(row[0] - np.max(tr)) <= 1e-9 * max(1.0, abs(np.max(tr)))
    assert max(ko.schedule_trace(m, types, times, bt["first_feasible_index"], P0, t_prev, t_end)) < thr
    lo_thr, _ = tree.score_schedules(types, times, t_prev, t_end, P0, threshold=thr - 0.5)
    hi_thr, _ = tree.score_schedules(types, times, t_prev, t_end, P0, threshold=thr + 0.5)
    assert lo_thr["n_feasible"] <= bt["n_feasible"] <= hi_thr["n_feasible"]
    assert lo_thr["argmin_sum_index"] == bt["argmin_sum_index"]  # the threshold does not enter the argmin


@pytest.mark.parametrize("L,n_low", [(18, 1), (18, 3), (20, 3), (10, 2), (6, 6)])
def test_strided_slices_are_bit_identical_rows_of_the_full_search(L, n_low, monkeypatch):
    """b200kf_score_schedules_strided (the sharded search's slices: the n_low earliest decisions fixed): every slice's
    per-schedule rows ARE the rows low | (j << n_low) of the full search, bit for bit, on the prefix-sharing kernel
    (L - n_low >= 14) and on the flat one, for ragged sub-ranges too; the slices' results merge (the reduce every rank
    applies after the all-gather) to the full search's answer."""
    from sensorfusion_kalmanfilter_b200.distributed import merge_best, shard_strided
    from sensorfusion_kalmanfilter_b200.engine import Engine
    m = ko.get_model("CA15")
    rng = np.random.default_rng(11)
    types = np.where(np.arange(L) % 5 == 1, 0, 1).astype(np.uint8)
    times = 1000.0 + np.cumsum(0.005 + 0.0005 * rng.random(L))
    P0 = m.P0() * 1e-8
    t_prev, t_end = 999.995, float(times[-1])
    eng = Engine("CA15")
    _, probe = eng.score_schedules(types, times, t_prev, t_end, P0, first=1, count=min(4096, (1 << L) - 1),
                                   want_per_schedule=True)
    thr = float(np.quantile(probe.cpu().numpy()[:, 0], 0.4))
    full, per = eng.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, want_per_schedule=True)
    per = per.cpu().numpy()
    parts = []
    for low in range(1 << n_low):
        if L > n_low:
            assert shard_strided(L, 1 << n_low, low) == (low, n_low)
        b, sl = eng.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, want_per_schedule=True,
                                    low_bits=low, n_low=n_low)
        sl = sl.cpu().numpy()
        idx = low | (np.arange(1 << (L - n_low), dtype=np.int64) << n_low)
        np.testing.assert_array_equal(sl, per[idx])
        parts.append(b)
    merged = merge_best(parts)
    for k in ("argmin_sum_index", "first_feasible_index", "n_feasible", "n_near_threshold", "argmin_sum_score"):
        assert merged[k] == full[k], k
    if L - n_low >= 8:  # a ragged sub-range of one slice (head and tail on the flat kernel around aligned tree blocks)
        low = (1 << n_low) - 1
        first, count = 37, (1 << (L - n_low)) - 37 - 101
        _, sl = eng.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, want_per_schedule=True,
                                    low_bits=low, n_low=n_low, first=first, count=count)
        idx = low | (np.arange(first, first + count, dtype=np.int64) << n_low)
        np.testing.assert_array_equal(sl.cpu().numpy(), per[idx])
    with pytest.raises(Exception):
        eng.score_schedules(types, times, t_prev, t_end, P0, low_bits=1 << n_low, n_low=n_low)
    with pytest.raises(Exception):
        eng.score_schedules(types, times, t_prev, t_end, P0, low_bits=0, n_low=n_low, first=0,
                            count=(1 << (L - n_low)) + 1)


@pytest.mark.parametrize("use_ld0", [False, True])
def test_decisions_taken_on_determinants_equal_the_decisions_on_the_logged_rows(use_ld0):
    """The search decides on the determinants (no logarithm per schedule: maxdet < exp(threshold), sum formed only for
    schedules that can win); the per-schedule rows carry the logarithms.  Recomputing winner, first feasible schedule
    and counts FROM THE ROWS must give the same answer (the two can differ only for schedules the call reports as
    undecided to rounding, n_near_threshold), with and without the caller's exact term 0."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    from sensorfusion_kalmanfilter_b200.kf_workers import _lex_key
    m = ko.get_model("CA15")
    L = 16
    rng = np.random.default_rng(5)
    types = np.where(np.arange(L) % 6 == 2, 0, 1).astype(np.uint8)
    times = 1000.0 + np.cumsum(0.005 + 0.0005 * rng.random(L))
    P0 = m.P0() * 1e-8
    t_prev, t_end = 999.995, float(times[-1])
    ld0 = float(np.linalg.slogdet(P0)[1]) if use_ld0 else None
    eng = Engine("CA15")
    _, per = eng.score_schedules(types, times, t_prev, t_end, P0, initial_logdet=ld0, want_per_schedule=True)
    rows = per.cpu().numpy()
    idx = np.arange(1, 1 << L)
    win = int(idx[np.lexsort((idx, rows[1:, 1]))[0]])
    for quant in (0.05, 0.4, 0.9):
        thr = float(np.quantile(rows[1:, 0], quant))
        for with_rows in (False, True):
            best, _ = eng.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, initial_logdet=ld0,
                                          want_per_schedule=with_rows)
            assert best["argmin_sum_index"] == win
            assert best["argmin_sum_score"] == rows[win, 1]
            feas = idx[rows[1:, 0] < thr]
            near = int(np.sum(np.abs(rows[1:, 0] - thr) <= 1e-9 * max(1.0, abs(thr))))
            if use_ld0 and ld0 >= thr:  # the caller's exact term 0 rules every schedule out on its own: none is undecided
                near = 0
            assert abs(best["n_near_threshold"] - near) <= 1  # the zone's two edges are themselves rounded
            assert abs(best["n_feasible"] - len(feas)) <= best["n_near_threshold"]
            if best["n_near_threshold"] == 0:
                assert best["n_feasible"] == len(feas)
                first = min(feas.tolist(), key=lambda s: _lex_key(s, L)) if len(feas) else None
                assert best["first_feasible_index"] == first
    # thresholds beyond any determinant: everything / nothing feasible, nothing near
    for thr, n in ((float("inf"), (1 << L) - 1), (800.0, (1 << L) - 1), (-800.0, 0), (float("-inf"), 0)):
        best, _ = eng.score_schedules(types, times, t_prev, t_end, P0, threshold=thr, initial_logdet=ld0)
        assert best["n_feasible"] == n and best["n_near_threshold"] == 0 and best["argmin_sum_index"] == win


# records of the 15- / 8-state models: 13 gain slots; one 3 x 3 and one 2 x 2 covariance (6 + 3 entries); the linear
# block's K is row-major [3][n_meas], the angular block's [2][2] starts at slot 9
_IR_IMU = {"CA15": ([0, 6, 12], [3, 9]), "CA8": ([0, 3, 6], [2, 5])}  # measured global states per block, in column order


def _gains_of_cov_records(fa, imu, name):
    """Gain slots [n, 13] of COVARIANCE records (pad = 1: g[0:9] = P+ of the two blocks, upper triangles): K = P+ H^T R^-1,
    the product the consumer forms (IMU: every state of a block measured; GPS: the first block's position only)."""
    from oracle import kf_oracle as ko
    m = ko.get_model(name)
    r_imu, r_gps = np.diag(m.R_imu()), np.diag(m.R_gps())
    lin_states, ang_states = _IR_IMU[name]
    # H_imu rows select states in the model's own order: the variance of measured STATE s is R_imu[row with H[row, s] = 1]
    H = m.H_imu()
    var_of_state = {int(np.argmax(H[r])): r_imu[r] for r in range(H.shape[0])}
    P = fa[:, 10:19]
    lin = np.zeros((len(fa), 3, 3))
    iu = np.triu_indices(3)
    lin[:, iu[0], iu[1]] = P[:, :6]
    lin = lin + np.swapaxes(lin, 1, 2) - np.einsum("fij,ij->fij", lin, np.eye(3))
    ang = np.zeros((len(fa), 2, 2))
    ju = np.triu_indices(2)
    ang[:, ju[0], ju[1]] = P[:, 6:9]
    ang = ang + np.swapaxes(ang, 1, 2) - np.einsum("fij,ij->fij", ang, np.eye(2))
    K = np.zeros((len(fa), 13))
    ir_lin = np.array([1.0 / var_of_state[s] for s in lin_states])
    ir_ang = np.array([1.0 / var_of_state[s] for s in ang_states])
    K[imu, 0:9] = (lin[imu] * ir_lin[None, None, :]).reshape(-1, 9)
    K[imu, 9:13] = (ang[imu] * ir_ang[None, None, :]).reshape(-1, 4)
    K[~imu, 0:3] = lin[~imu][:, :, 0] / r_gps[0]
    return K


def _check_gain_records(ra, rb, rtol, name="CA15"):
    """Records [n, bytes] (uint8) of two producers: (dt, det[4], ldet[4], (type, pad), 13 slots).  pad = 0: GAIN records
    (slots = K, det[0] = det P); pad = 1: COVARIANCE records (slots 0..8 = P+, nothing else written).  dt and the event
    type bit for bit; two records of the same kind slot by slot to `rtol` (0 = bit for bit; an IMU event writes every
    gain slot, a GPS event only the first block's n x 1 column -- unwritten slots are not compared); a covariance record
    against a gain record through K = P+ H^T R^-1."""
    fa = np.ascontiguousarray(ra).view(np.float64)
    fb = np.ascontiguousarray(rb).view(np.float64)
    ia = np.ascontiguousarray(ra).view(np.int32).reshape(len(ra), -1)
    ib = np.ascontiguousarray(rb).view(np.int32).reshape(len(rb), -1)
    np.testing.assert_array_equal(ia[:, 18], ib[:, 18])  # event type (byte offset 72)
    np.testing.assert_array_equal(fa[:, 0], fb[:, 0])    # dt
    if len(fa) == 0:
        return
    imu = ia[:, 18] == 1
    cov_a, cov_b = bool((ia[:, 19] == 1).all()), bool((ib[:, 19] == 1).all())
    assert cov_a or bool((ia[:, 19] == 0).all())
    assert cov_b or bool((ib[:, 19] == 0).all())
    if cov_a and cov_b:
        cmp = [(slice(10, 19), slice(None))]
    elif not cov_a and not cov_b:
        cmp = [(slice(1, 5), slice(None)), (slice(10, 23), imu), (slice(10, 13), ~imu)]
    else:
        ka = _gains_of_cov_records(fa, imu, name) if cov_a else fa[:, 10:23]
        kb = _gains_of_cov_records(fb, imu, name) if cov_b else fb[:, 10:23]
        np.testing.assert_allclose(ka[imu], kb[imu], rtol=max(rtol, 1e-13), atol=0)
        np.testing.assert_allclose(ka[~imu, :3], kb[~imu, :3], rtol=max(rtol, 1e-13), atol=0)
        return
    for sl, rows in cmp:
        if rtol == 0:
            np.testing.assert_array_equal(fa[rows, sl], fb[rows, sl])
        else:
            np.testing.assert_allclose(fa[rows, sl], fb[rows, sl], rtol=rtol, atol=0)


@pytest.mark.parametrize("name", ["CA15", "CA8"])
def test_warp_cooperative_selection_is_bit_identical_to_the_thread_walk(name, monkeypatch):
    """b200kf_select_greedy with few instances runs one WARP per instance (events loaded 32 at a time, window
    closing found with ballots); B200KF_SELECT_WARP=0 keeps the thread-per-instance walk.  Same selections and
    covariances bit for bit, the warp walk's covariance records give the thread walk's gain records, over a stream with batch-straddling queues, empty queues
    (period shorter than the event spacing) and long queues (period of several batches)."""
    from sensorfusion_kalmanfilter_b200 import streams, synth
    from sensorfusion_kalmanfilter_b200.engine import Engine
    ev = synth.make_indexed_sensor_data(5003, seed=8)[3:]
    ps = streams.PackedStream.from_events(ev)
    types, times = ps.types[1:], ps.times[1:]
    periods = [1.0 / f for f in (3.0, 10.0, 55.0, 120.0, 180.0, 1000.0)]
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("B200KF_SELECT_WARP", flag)
        eng = Engine(name)
        P0 = np.repeat(eng.spec.default_P0()[None], len(periods), axis=0)
        for mode in ("symmetric", "decoupled", "dense"):
            want_gains = mode == "symmetric"
            got = eng.select_greedy(types, times, periods, [float(ps.times[0])] * len(periods), P0, mode=mode,
                                    want_gains=want_gains)
            outs[(flag, mode)] = got
    for mode in ("symmetric", "decoupled", "dense"):
        a, b = outs[("1", mode)], outs[("0", mode)]
        for sa, sb in zip(a[0], b[0]):
            np.testing.assert_array_equal(sa, sb)
        assert bool((a[1] == b[1]).all())
        if mode == "symmetric":
            ra, rb = a[2].cpu().numpy(), b[2].cpu().numpy()
            for i, sa in enumerate(a[0]):  # (the two kernels may contract the determinant's FMAs differently)
                _check_gain_records(ra[i, :len(sa)], rb[i, :len(sa)], rtol=1e-13, name=name)
    # and the selections are the oracle's (greedy_schedule over the window queues, kf_workers.py:877-910)
    if name == "CA15":
        st, _, _ = ko.run_scheduled(ko.get_model(name), ev, 55.0, 0, len(ev))
        sel_ref = np.searchsorted(ps.times, np.array([s[0] for s in st[1:]])) - 1
        np.testing.assert_array_equal(outs[("1", "symmetric")][0][2], sel_ref)


def test_chunked_selection_equals_one_walk_bit_for_bit():
    """b200kf_select_greedy_chunk: the greedy walk cut at arbitrary event indices (mid-queue, mid-batch, empty and
    one-event chunks) gives the selections, covariances and gain records of ONE walk over the stream."""
    import torch
    from sensorfusion_kalmanfilter_b200 import streams, synth
    from sensorfusion_kalmanfilter_b200.engine import Engine
    ev = synth.make_indexed_sensor_data(6003, seed=12)[3:]
    ps = streams.PackedStream.from_events(ev)
    types, times = ps.types[1:], ps.times[1:]
    E = len(types)
    periods = [1.0 / f for f in (3.0, 10.0, 55.0, 120.0, 1000.0)]
    eng = Engine("CA15")
    P0 = np.repeat(eng.spec.default_P0()[None], len(periods), axis=0)
    sels, P_fin, recs = eng.select_greedy(types, times, periods, [float(ps.times[0])] * len(periods), P0,
                                          mode="symmetric", want_gains=True)
    st = eng.select_state(types, times, periods, float(ps.times[0]), P0)
    cuts = [1, 2, 33, 64, 64, 1000, 1001, 2500, 4097, E - 1, E]
    for c in cuts:
        st.advance(c)
    torch.cuda.synchronize()
    ns = st.n_sel.cpu().numpy()
    for i, sa in enumerate(sels):
        assert ns[i] == len(sa)
        np.testing.assert_array_equal(st.sel[i, :len(sa)].cpu().numpy(), sa)
        _check_gain_records(st.recs[i, :len(sa)].cpu().numpy(), recs[i, :len(sa)].cpu().numpy(), rtol=0)
    assert bool((st.P == P_fin).all())
