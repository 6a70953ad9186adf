"""Parity at the LENGTHS BASELINE.json names, not just at a few hundred
steps: a rounding difference between the CUDA formulation (sequential scalar / closed-form updates,
R-normalised lean step, running-product log det) and the reference's `inv`-based numpy form would have
10 000 - 590 000 steps to grow here.

* config 3: CV6, a 131 072-wide launch with one covariance per filter x 10 000 steps; 64 random filters are
  copied back and replayed through the oracle over ALL steps (x, P, max / sum / last log det at 1e-9);
* config 5: BIAS9, 32 filters of an 8 192-wide launch x 50 000 steps, FP64 at 1e-9 and FP32 at 1e-4, both
  against the ORACLE (not against the other GPU path);
* config 2: the scheduled filter (kf_workers.py:826-957) over the whole 2 808 s stream at 10 Hz and 120 Hz for
  four Monte-Carlo seeds against golden vectors written by the UNMODIFIED reference
  (oracle/make_golden_long.py), through the batched sweep (with and without precomputed gains) and through
  the reference-facing method; plus run_kalman_filter_full over 120 000 events.
"""
import os

import numpy as np
import pytest

from oracle import kf_oracle as ko
from sensorfusion_kalmanfilter_b200 import streams, synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def agg_err(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


# ----------------------------------------------------------------------------------------------
# config 3: CV6 x 10 000 steps
# ----------------------------------------------------------------------------------------------
def test_cv6_config3_length_64_filters_of_a_full_width_launch_vs_oracle():
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine, fill_payload
    N, T, S = 131072, 10000, 64
    eng = Engine("CV6")
    dev = eng.device
    types, dts = synth.gps_imu_schedule(T)
    types_d = torch.from_numpy(types).to(dev)
    pay = fill_payload(eng, types_d, N, seed=77)
    rng = np.random.default_rng(3)
    scale = 1.0 + rng.random(N)  # every filter its own covariance: the per-filter ring kernel
    P0 = eng.spec.default_P0()[None] * scale[:, None, None]
    x0 = rng.normal(0.0, 1.0, (N, 6))
    out = eng.run_batch(types_d, torch.from_numpy(dts).to(dev), pay, x0, P0)
    assert int(out["status"].sum()) == 0
    pick = np.sort(rng.choice(N, S, replace=False))
    pick_d = torch.from_numpy(pick).to(dev)
    pay_h = pay.index_select(2, pick_d).cpu().numpy()
    ref = ko.run_batch(ko.get_model("CV6"), types, dts, pay_h, x0[pick], P0[pick], want_traj=False)
    assert relmax(out["x"][pick_d].cpu().numpy(), ref["x"]) < 1e-9
    assert relmax(out["P"][pick_d].cpu().numpy(), ref["P"]) < 1e-9
    agg = out["agg"][pick_d].cpu().numpy()
    assert agg_err(agg[:, :3], ref["agg"]) < 1e-9
    assert np.all(agg[:, 3] == T)
    # the same 64 filters as a launch of their own (another CTA shape, LDG or ring) agree with the wide launch
    small = eng.run_batch(types, dts, pay_h, x0[pick], P0[pick])
    assert relmax(small["x"].cpu().numpy(), ref["x"]) < 1e-9
    assert relmax(small["P"].cpu().numpy(), ref["P"]) < 1e-9
    # trajectory mode over the whole length: the per-step outputs of a few filters against the oracle's
    reft = ko.run_batch(ko.get_model("CV6"), types, dts, pay_h[:, :, :8], x0[pick[:8]], P0[pick[:8]])
    outt = eng.run_batch(types, dts, pay_h, x0[pick], P0[pick], traj_ncomp=6, want_logdet=True, want_agg=False)
    assert relmax(outt["traj"][:, :, :8].cpu().numpy(), reft["traj"]) < 1e-9
    ld = outt["logdet"][:, :8].cpu().numpy()
    assert agg_err(ld, reft["logdet"]) < 1e-9


# ----------------------------------------------------------------------------------------------
# config 5: BIAS9 x 50 000 steps, both precisions against the oracle
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def bias9_case():
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine, fill_payload
    N, T, S = 8192, 50000, 32
    eng = Engine("BIAS9")
    types, dts = synth.gps_imu_schedule(T)
    pay = fill_payload(eng, torch.from_numpy(types).to(eng.device), N, seed=99)
    rng = np.random.default_rng(4)
    scale = 1.0 + rng.random(N)
    P0 = eng.spec.default_P0()[None] * scale[:, None, None]
    x0 = rng.normal(0.0, 1.0, (N, 9))
    pick = np.sort(rng.choice(N, S, replace=False))
    pay_h = pay.index_select(2, torch.from_numpy(pick).to(eng.device)).cpu().numpy()
    ref = ko.run_batch(ko.get_model("BIAS9"), types, dts, pay_h, x0[pick], P0[pick], want_traj=False)
    return types, dts, pay, x0, P0, pick, ref


@pytest.mark.parametrize("precision,tol", [("f64", 1e-9), ("f32", 1e-4)])
def test_bias9_config5_length_vs_oracle(bias9_case, precision, tol):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    types, dts, pay, x0, P0, pick, ref = bias9_case
    T = len(types)
    eng = Engine("BIAS9", precision=precision)
    out = eng.run_batch(types, dts, pay, x0, P0)
    assert int(out["status"].sum()) == 0
    assert relmax(out["x"].cpu().numpy()[pick], ref["x"]) < tol
    assert relmax(out["P"].cpu().numpy()[pick], ref["P"]) < tol
    agg = out["agg"].cpu().numpy()[pick]
    assert agg_err(agg[:, :3], ref["agg"]) < tol
    assert np.all(agg[:, 3] == T)


# ----------------------------------------------------------------------------------------------
# config 2 at full length against the unmodified reference
# ----------------------------------------------------------------------------------------------
def _checksum(ty, tm, da):
    return np.array([float(ty.sum()), float(tm.sum()), float(np.abs(da).sum()), float((da * da).sum())])


def _sample(a, stride):
    a = np.asarray(a)
    return np.concatenate([a[:-8:stride], a[-8:]], axis=0)


@pytest.fixture(scope="module")
def config2():
    g = np.load(os.path.join(GOLDEN, "ca15_config2_full_length.npz"))
    n = int(g["n_events"])
    ev = synth.make_indexed_sensor_data(n + 3, seed=int(g["base_seed"]))[3:]
    ps = streams.PackedStream.from_events(ev)
    # the inputs are regenerated, not stored: a generator drift must not masquerade as a parity failure
    np.testing.assert_array_equal(_checksum(ps.types, ps.times, ps.data), g["checksum"])
    noise = [synth.seed_noise(ps.types, int(s)) for s in g["seeds"]]
    np.testing.assert_array_equal(np.array([float(np.abs(z).sum()) for z in noise]), g["noise_checksum"])
    return g, ev, ps, noise


@pytest.mark.parametrize("use_gains", [False, True])
def test_config2_full_length_sweep_vs_reference(config2, use_gains):
    """12-rate sweep machinery (scheduled_sweep) on the two rates the golden file holds, four seeds each."""
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine
    g, ev, ps, noise = config2
    stride = int(g["stride"])
    rates = [int(r) for r in g["rates"]]
    seeds = [int(s) for s in g["seeds"]]
    eng = Engine("CA15")
    data = np.stack([ps.data + z for z in noise], axis=2)  # [E, 9, seeds]
    pay = torch.from_numpy(np.ascontiguousarray(data[1:])).to(eng.device)
    x0 = np.zeros((len(seeds), 15))
    x0[:, :3] = data[0, :3, :].T  # first GPS fix initialises the position (kf_workers.py:854-862)
    out = eng.scheduled_sweep(ps.types[1:], ps.times[1:], pay, rates, x0, eng.spec.default_P0(), float(ps.times[0]),
                              traj_ncomp=6, use_gains=use_gains)
    for f in rates:
        sel_ref = np.cumsum(g[f"r{f}_sel_delta"].astype(np.int64))
        r = out[f]
        np.testing.assert_array_equal(np.asarray(r["sel"], dtype=np.int64) + 1, sel_ref)  # bit-exact selection
        traj = r["traj"].cpu().numpy()  # [T, 6, seeds]
        times = r["times"].cpu().numpy()
        for j, s in enumerate(seeds):
            st_ref = g[f"r{f}_s{s}_states"]
            st = np.concatenate([[np.concatenate([[ps.times[0]], x0[j, :6]])],
                                 np.concatenate([times[:, None], traj[:, :, j]], axis=1)], axis=0)
            assert len(st) == int(g[f"r{f}_s{s}_n_states"])
            got = _sample(st, stride)
            np.testing.assert_array_equal(got[:, 0], st_ref[:, 0])
            assert relmax(got[:, 1:], st_ref[:, 1:]) < 1e-9
            assert relmax(r["P"][j].cpu().numpy(), g[f"r{f}_s{s}_P"]) < 1e-9
            a = r["agg"][j].cpu().numpy()
            ref_agg = g[f"r{f}_s{s}_agg"].copy()
            ld0 = np.linalg.slogdet(eng.spec.default_P0())[1]  # the reference's list starts with slogdet(P0) (:869-870)
            ref_wo0 = np.array([ref_agg[0], ref_agg[1] - ld0, ref_agg[2]])
            assert abs(a[1] - ref_wo0[1]) < 1e-9 * abs(ref_wo0[1])
            assert abs(a[2] - ref_wo0[2]) < 1e-9 * max(1.0, abs(ref_wo0[2]))
            assert abs(max(a[0], ld0) - ref_agg[0]) < 1e-9 * max(1.0, abs(ref_agg[0]))


@pytest.mark.parametrize("use_gains", [False, True])
def test_config2_full_length_time_chunked_sweep_vs_reference(config2, use_gains):
    """The same sweep in time chunks with state carried between them (how 4096 seeds x 589 679 events, 174 GB of
    payload, run on one GPU): final covariance and the log-det aggregates against the reference, and the
    final state against the one-launch sweep."""
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine
    g, ev, ps, noise = config2
    rates = [int(r) for r in g["rates"]]
    seeds = [int(s) for s in g["seeds"]]
    eng = Engine("CA15")
    data = np.stack([ps.data + z for z in noise], axis=2)
    pay = torch.from_numpy(np.ascontiguousarray(data[1:])).to(eng.device)
    x0 = np.zeros((len(seeds), 15))
    x0[:, :3] = data[0, :3, :].T
    P0 = eng.spec.default_P0()
    one = eng.scheduled_sweep(ps.types[1:], ps.times[1:], pay, rates, x0, P0, float(ps.times[0]), use_gains=use_gains)
    out = eng.scheduled_sweep_chunked(ps.types[1:], ps.times[1:], rates, x0, P0, float(ps.times[0]),
                                      lambda e0, e1: pay[e0:e1], 70001, use_gains=use_gains)
    ld0 = np.linalg.slogdet(P0)[1]
    for f in rates:
        np.testing.assert_array_equal(np.asarray(out[f]["sel"], dtype=np.int64) + 1,
                                      np.cumsum(g[f"r{f}_sel_delta"].astype(np.int64)))
        assert relmax(out[f]["x"].cpu().numpy(), one[f]["x"].cpu().numpy()) < 1e-12
        for j, s in enumerate(seeds):
            assert relmax(out[f]["P"][j].cpu().numpy(), g[f"r{f}_s{s}_P"]) < 1e-9
            a, ref_agg = out[f]["agg"][j].cpu().numpy(), g[f"r{f}_s{s}_agg"]
            assert abs(a[1] - (ref_agg[1] - ld0)) < 1e-9 * abs(ref_agg[1] - ld0)
            assert abs(a[2] - ref_agg[2]) < 1e-9 * max(1.0, abs(ref_agg[2]))
            assert a[3] == int(g[f"r{f}_s{s}_n_states"]) - 1


def test_config2_full_length_pipelined_sweep_equals_chunked_sweep(config2):
    """The pipelined sweep (selection walked in chunks one chunk ahead of the filters, on its own stream) against
    the chunked sweep with precomputed gains: same selections, final states, covariances and aggregates."""
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine
    g, ev, ps, noise = config2
    rates = [int(r) for r in g["rates"]] + [50]
    eng = Engine("CA15")
    data = np.stack([ps.data + z for z in noise], axis=2)
    pay = torch.from_numpy(np.ascontiguousarray(data[1:])).to(eng.device)
    x0 = np.zeros((data.shape[2], 15))
    x0[:, :3] = data[0, :3, :].T
    P0 = eng.spec.default_P0()
    args = (ps.types[1:], ps.times[1:], rates, x0, P0, float(ps.times[0]), lambda e0, e1: pay[e0:e1], 60007)
    ref = eng.scheduled_sweep_chunked(*args, use_gains=True)

    def guarded_chunk(e0, e1):  # a copy with NaN rows around it: a read outside [e0, e1) poisons the state
        buf = torch.full((e1 - e0 + 8, 9, pay.shape[2]), float("nan"), dtype=pay.dtype, device=pay.device)
        buf[4:-4] = pay[e0:e1]
        return buf[4:-4]

    # payload_depth=3: the callback cycles through three buffers (each refilled, NaN around the rows, when its turn
    # comes again) and the rates' streams meet only where a buffer is reused
    pool = [torch.empty((60007 + 1024, 9, pay.shape[2]), dtype=pay.dtype, device=pay.device) for _ in range(3)]
    turn = {"c": 0}

    def cycling_chunk(e0, e1):
        buf = pool[turn["c"] % 3]
        turn["c"] += 1
        buf.fill_(float("nan"))
        buf[4:4 + e1 - e0] = pay[e0:e1]
        return buf[4:4 + e1 - e0]

    for out in (eng.scheduled_sweep_pipelined(*args[:6], guarded_chunk, 60007),
                eng.scheduled_sweep_pipelined(*args[:6], cycling_chunk, 60007, payload_depth=3)):
        for f in rates:
            np.testing.assert_array_equal(out[f]["sel"].cpu().numpy(), np.asarray(ref[f]["sel"]))
            assert bool((out[f]["x"] == ref[f]["x"]).all())
            assert bool((out[f]["P"] == ref[f]["P"]).all())
            a, b = out[f]["agg"].cpu().numpy(), ref[f]["agg"].cpu().numpy()
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    for f in (10, 120):
        np.testing.assert_array_equal(out[f]["sel"].cpu().numpy().astype(np.int64) + 1,
                                      np.cumsum(g[f"r{f}_sel_delta"].astype(np.int64)))


@pytest.mark.parametrize("rate,seed", [(10, 1), (120, 2)])
def test_config2_full_length_reference_facing_method(config2, rate, seed):
    from oracle.make_golden import arrays_to_events
    from sensorfusion_kalmanfilter_b200.kf_workers import KF_SensorFusion
    g, ev, ps, noise = config2
    stride = int(g["stride"])
    j = [int(s) for s in g["seeds"]].index(seed)
    k = KF_SensorFusion("gps_data.csv", "imu_data.csv")
    k.indexed_sensor_data = arrays_to_events(ps.types, ps.times, ps.data + noise[j])
    k.set_processing_frequency(rate)
    st, ld, P = k.run_kalman_filter_scheduled(0, len(ev), None, None, "greedy")
    assert len(st) == int(g[f"r{rate}_s{seed}_n_states"])
    st = _sample(np.array([[float(v) for v in s] for s in st]), stride)
    ref = g[f"r{rate}_s{seed}_states"]
    np.testing.assert_array_equal(st[:, 0], ref[:, 0])
    assert relmax(st[:, 1:], ref[:, 1:]) < 1e-9
    ldr = g[f"r{rate}_s{seed}_logdet"]
    assert agg_err(_sample(np.array(ld, dtype=np.float64), stride), ldr) < 1e-9
    assert relmax(P, g[f"r{rate}_s{seed}_P"]) < 1e-9


def test_full_filter_120k_events_vs_reference():
    from sensorfusion_kalmanfilter_b200.kf_workers import KF_SensorFusion
    g = np.load(os.path.join(GOLDEN, "ca15_full_120k.npz"))
    n, stride = int(g["n_events"]), int(g["stride"])
    ev = synth.make_indexed_sensor_data(n + 3, seed=int(g["base_seed"]))[3:]
    ps = streams.PackedStream.from_events(ev)
    np.testing.assert_array_equal(_checksum(ps.types, ps.times, ps.data), g["checksum"])
    k = KF_SensorFusion("gps_data.csv", "imu_data.csv")
    k.indexed_sensor_data = ev
    st, ld, P, prev = k.run_kalman_filter_full(0, len(ev))
    assert len(st) == int(g["n_states"]) and prev == float(g["prev_time"])
    st = _sample(np.array([[float(v) for v in s] for s in st]), stride)
    np.testing.assert_array_equal(st[:, 0], g["states"][:, 0])
    assert relmax(st[:, 1:], g["states"][:, 1:]) < 1e-9
    lda = np.array(ld, dtype=np.float64)
    assert agg_err(_sample(lda, stride), g["logdet"]) < 1e-9
    assert agg_err(np.array([lda.max(), lda.sum(), lda[-1]]), g["agg"]) < 1e-9
    assert relmax(P, g["P"]) < 1e-9
