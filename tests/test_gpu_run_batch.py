"""GPU parity of b200kf_run_batch (through the C ABI) against the numpy oracle on identical seeded
inputs.  Tolerances are BASELINE.json's: FP64 x, P within 1e-9 relative, log det within
1e-9*max(1,|logdet|); FP32 mode 1e-4."""
import numpy as np
import pytest

from oracle import kf_oracle as ko
from sensorfusion_kalmanfilter_b200 import synth

pytestmark = pytest.mark.gpu

RTOL64 = 1e-9
RTOL32 = 1e-4


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def make_case(name, N, T, seed, dense_P0=False):
    m = ko.get_model(name)
    rng = np.random.default_rng(seed)
    types, dts = synth.gps_imu_schedule(T, gps_every=7, dt=0.005)
    dts = dts * (1.0 + 0.3 * rng.random(T))
    dts[0] = 0.0
    payload = synth.make_payload(m.d, types, N, seed=seed, m_gps=m.m_gps)
    x0 = rng.normal(0, 1.0, (N, m.n))
    P0 = np.repeat(m.P0()[None], N, axis=0)
    if dense_P0:
        A = rng.normal(size=(N, m.n, m.n))
        P0 = A @ np.swapaxes(A, 1, 2) + 5.0 * np.eye(m.n)
    return m, types, dts, payload, x0, P0


@pytest.mark.parametrize("name", ["CV6", "CA8", "BIAS9", "CA15", "CVBU6"])
@pytest.mark.parametrize("mode", ["decoupled", "dense", "auto", "symmetric", "shared"])
def test_run_batch_f64_matches_oracle(name, mode):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 70, 160
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=3)
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    eng = Engine(name)
    out = eng.run_batch(types, dts, payload, x0, P0, traj_ncomp=m.n, want_logdet=True, mode=mode)
    assert int(out["status"].sum()) == 0
    assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL64
    assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL64
    assert rel(out["traj"].cpu().numpy(), ref["traj"]) < RTOL64
    ld = out["logdet"].cpu().numpy()
    assert np.max(np.abs(ld - ref["logdet"]) / np.maximum(1.0, np.abs(ref["logdet"]))) < RTOL64
    agg = out["agg"].cpu().numpy()
    assert np.max(np.abs(agg[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL64
    assert np.all(agg[:, 3] == T)
    # aggregate-only launch (the throughput path) gives the same final state and aggregates
    out2 = eng.run_batch(types, dts, payload, x0, P0, mode=mode)
    np.testing.assert_allclose(out2["x"].cpu().numpy(), out["x"].cpu().numpy(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(out2["agg"].cpu().numpy(), agg, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["CV6", "CA15", "BIAS9", "CVBU6"])
def test_dense_P0_uses_dense_kernel_and_matches(name):
    """An arbitrary dense initial covariance must take the dense path in auto mode."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 40, 90
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=9, dense_P0=True)
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    eng = Engine(name)
    out = eng.run_batch(types, dts, payload, x0, P0, want_logdet=True, mode="auto")
    assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL64
    assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL64
    ld = out["logdet"].cpu().numpy()
    assert np.max(np.abs(ld - ref["logdet"]) / np.maximum(1.0, np.abs(ref["logdet"]))) < RTOL64
    # one coupled filter among block-diagonal ones still forces the dense path for the launch
    P1 = np.repeat(m.P0()[None], N, axis=0)
    P1[N // 2] = P0[N // 2]
    ref1 = ko.run_batch(m, types, dts, payload, x0, P1)
    out1 = eng.run_batch(types, dts, payload, x0, P1, mode="auto")
    assert rel(out1["P"].cpu().numpy(), ref1["P"]) < RTOL64


@pytest.mark.parametrize("name", ["CV6", "BIAS9", "CA15", "CVBU6"])
def test_run_batch_f32_mode(name):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 64, 300
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=5)
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    eng = Engine(name, precision="f32")
    out = eng.run_batch(types, dts, payload, x0, P0, want_logdet=True)
    assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL32
    assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL32
    ld = out["logdet"].cpu().numpy()
    assert np.max(np.abs(ld - ref["logdet"]) / np.maximum(1.0, np.abs(ref["logdet"]))) < RTOL32


@pytest.mark.parametrize("name", ["CV6", "BIAS9", "CVBU6"])
def test_f32_lean_aggregate_over_a_long_launch(name):
    """The float32 ring kernels keep the log-det aggregate in float32 (running mantissa product + integer
    exponent sum): over 30 000 steps max / sum / last must still agree with the float64 path (itself checked
    against the oracle above) to the float32 tolerance, with per-filter and with identical covariances."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 64, 30000
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=17)
    scale = 1.0 + (np.arange(N) % 16) / 16.0
    for P in (P0 * scale[:, None, None], P0):
        ref = Engine(name).run_batch(types, dts, payload, x0, P, mode="symmetric")
        for mode in ("symmetric", "auto"):
            out = Engine(name, precision="f32").run_batch(types, dts, payload, x0, P, mode=mode)
            assert int(out["status"].sum()) == 0
            assert rel(out["x"].cpu().numpy(), ref["x"].cpu().numpy()) < RTOL32
            assert rel(out["P"].cpu().numpy(), ref["P"].cpu().numpy()) < RTOL32
            a, b = out["agg"].cpu().numpy(), ref["agg"].cpu().numpy()
            assert np.max(np.abs(a[:, :3] - b[:, :3]) / np.maximum(1.0, np.abs(b[:, :3]))) < RTOL32
            assert np.all(a[:, 3] == T)


def test_chunked_launches_continue_state_and_aggregates():
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 33, 240
    m, types, dts, payload, x0, P0 = make_case("CV6", N, T, seed=7)
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    eng = Engine("CV6")
    x, P, agg = x0, P0, None
    for k0 in range(0, T, 100):
        sl = slice(k0, min(T, k0 + 100))
        out = eng.run_batch(types[sl], dts[sl], payload[sl], x, P, agg=agg)
        x, P, agg = out["x"], out["P"], out["agg"]
    assert rel(x.cpu().numpy(), ref["x"]) < RTOL64
    assert rel(P.cpu().numpy(), ref["P"]) < RTOL64
    a = agg.cpu().numpy()
    assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL64
    assert np.all(a[:, 3] == T)


def test_gate_and_predict_only():
    """Adaptive gate (kf_workers.py:1023-1025) and predict-only events (:1060-1159)."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 48, 120
    m, types, dts, payload, x0, P0 = make_case("CA15", N, T, seed=13)
    base = ko.run_batch(m, types, dts, payload, x0, P0)
    thr = float(np.median(base["logdet"][20:])) + 2.0
    ref = ko.run_batch(m, types, dts, payload, x0, P0, gate_threshold=thr)
    eng = Engine("CA15")
    out = eng.run_batch(types, dts, payload, x0, P0, gate_threshold=thr, want_used=True, want_logdet=True)
    used = out["used"].cpu().numpy().astype(bool)
    # A decision may differ from the oracle's only where the oracle's predicted log det sits within
    # EPS of the threshold (a rounding difference could flip it); anywhere else a difference is a failure,
    # and so is any difference in x / P of a filter whose decisions all agree.
    EPS = 1e-9 * max(1.0, abs(thr))
    xs, Ps = x0.copy(), P0.copy()
    margin = np.full((T, N), np.inf)
    for k in range(T):  # the oracle's gate quantity, step by step (kf_workers.py:1023-1025)
        xp, Pp = ko.predict(m, xs, Ps, float(dts[k]))
        s, l = ko.logdet(Pp)
        margin[k] = np.abs(s * l - thr)
        xu, Pu = ko.step(m, xs, Ps, int(types[k]), float(dts[k]), np.ascontiguousarray(payload[k].T))
        g = ref["used"][k]
        xs = np.where(g[:, None], xu, xp)
        Ps = np.where(g[:, None, None], Pu, Pp)
    differ = used != ref["used"]
    for f in np.flatnonzero(differ.any(axis=0)):  # after its first flip a filter follows another path
        k_first = int(np.argmax(differ[:, f]))
        assert margin[k_first, f] <= EPS, "gate decision differs outside EPS of the threshold"
    same = ~differ.any(axis=0)  # filters with identical decisions throughout
    assert same.sum() >= N - 1  # at most a stray borderline filter, never the bulk
    assert rel(out["x"].cpu().numpy()[same], ref["x"][same]) < RTOL64
    assert rel(out["P"].cpu().numpy()[same], ref["P"][same]) < RTOL64
    assert 0.05 < used.mean() < 0.95
    types2 = np.full(T, 2, dtype=np.uint8)
    xs, Ps = x0.copy(), P0.copy()
    for k in range(T):
        xs, Ps = ko.predict(m, xs, Ps, float(dts[k]))
    out2 = eng.run_batch(types2, dts, payload, x0, P0)
    assert rel(out2["x"].cpu().numpy(), xs) < RTOL64
    assert rel(out2["P"].cpu().numpy(), Ps) < RTOL64


def test_host_buffer_path_matches_device_path():
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 257, 500
    m, types, dts, payload, x0, P0 = make_case("CV6", N, T, seed=21)
    eng = Engine("CV6")
    dev = eng.run_batch(types, dts, payload, x0, P0)
    host = eng.run_batch_host(types, dts, payload, x0, P0, chunk_steps=64)
    np.testing.assert_allclose(host["x"], dev["x"].cpu().numpy(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(host["P"], dev["P"].cpu().numpy(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(host["agg"], dev["agg"].cpu().numpy(), rtol=1e-12, atol=1e-12)
    assert host["status"].sum() == 0


def test_pageable_payload_is_staged_on_host_threads_and_matches_the_pinned_path():
    """A drop-in caller's payload is a plain numpy array (pageable): b200kf_run_batch_host stages it through two pinned
    chunk buffers filled by several host threads (chunks above 8 MB are split into per-thread slices).  Same bytes
    reach the device, so the results equal the pinned-buffer call's bit for bit -- over several chunks with a ragged
    last one, and a second call reusing the staging buffers."""
    import torch
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 16384, 100
    m, types, dts, payload, x0, P0 = make_case("CV6", N, T, seed=33)
    eng = Engine("CV6")
    pinned = torch.from_numpy(np.ascontiguousarray(payload)).pin_memory()
    ref = eng.run_batch_host(types, dts, pinned, x0, P0, chunk_steps=32)
    for rep in range(2):
        got = eng.run_batch_host(types, dts, np.ascontiguousarray(payload).copy(), x0, P0, chunk_steps=32)
        for k in ("x", "P", "agg", "status"):
            np.testing.assert_array_equal(got[k], ref[k])
    dev = eng.run_batch(types, dts, payload, x0, P0)
    np.testing.assert_allclose(ref["x"], dev["x"].cpu().numpy(), rtol=1e-13, atol=1e-13)


def test_ragged_and_empty_inputs():
    from sensorfusion_kalmanfilter_b200.engine import Engine
    eng = Engine("CV6")
    m = ko.get_model("CV6")
    # N = 1 filter, T = 1 step
    mm, types, dts, payload, x0, P0 = make_case("CV6", 1, 1, seed=2)
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    out = eng.run_batch(types, dts, payload, x0, P0)
    assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL64
    # T = 0: state passes through unchanged
    out0 = eng.run_batch(np.zeros(0, np.uint8), np.zeros(0), np.zeros((0, 3, 1)), x0, P0)
    np.testing.assert_array_equal(out0["x"].cpu().numpy(), x0)
    np.testing.assert_allclose(out0["P"].cpu().numpy(), P0)
    # non-multiple-of-block filter count
    mm, types, dts, payload, x0, P0 = make_case("CV6", 131, 37, seed=4)
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    out = eng.run_batch(types, dts, payload, x0, P0)
    assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL64
    with pytest.raises(ValueError):
        eng.run_batch(types, dts, payload[:, :2], x0, P0)


@pytest.mark.parametrize("name", ["CA15", "CV6"])
def test_one_gate_threshold_per_filter(name):
    """b200kf_run_batch_gated: filters of one launch with different thresholds (the reference's experiment runs
    the adaptive filter once per window, each with its own R_threshold, kf_workers.py:2316) equal the same
    filters launched one threshold at a time, and the oracle."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 12, 220
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=23)
    eng = Engine(name)
    full = eng.run_batch(types, dts, payload, x0, P0, want_logdet=True)
    ld = full["logdet"].cpu().numpy()
    qs = float(np.median(ld[T // 2:])) + np.array([0.5, 1.0, 2.0, 4.0])  # above the steady state: updates are skipped
    thr = np.concatenate([np.repeat(qs, 3)[: N - 1], [-1.0e300]])  # the last filter always updates
    out = eng.run_batch(types, dts, payload, x0, P0, gate_threshold=thr, want_used=True, want_logdet=True)
    assert int(out["status"].sum()) == 0
    for f in range(N):
        one = eng.run_batch(types, dts, payload[:, :, f:f + 1], x0[f:f + 1], P0[f:f + 1], gate_threshold=float(thr[f]),
                            want_used=True, want_logdet=True)
        np.testing.assert_array_equal(out["used"].cpu().numpy()[:, f], one["used"].cpu().numpy()[:, 0])
        assert rel(out["x"].cpu().numpy()[f], one["x"].cpu().numpy()[0]) < RTOL64
        assert rel(out["P"].cpu().numpy()[f], one["P"].cpu().numpy()[0]) < RTOL64
    ref = ko.run_batch(m, types, dts, payload[:, :, :1], x0[:1], P0[:1], gate_threshold=float(thr[0]))
    assert rel(out["x"].cpu().numpy()[0], ref["x"][0]) < RTOL64
    assert rel(out["P"].cpu().numpy()[0], ref["P"][0]) < RTOL64
    u = out["used"].cpu().numpy()
    assert u[:, -1].all() and all(0 < u[:, f].sum() < T for f in range(N - 1))
    assert len({int(u[:, f].sum()) for f in range(0, N - 1, 3)}) > 1  # different thresholds, different decisions
    assert rel(out["x"].cpu().numpy()[-1], full["x"].cpu().numpy()[-1]) < RTOL64


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_tma_staged_kernel_matches_ldg_kernel_and_oracle(precision, monkeypatch):
    """The aggregate-mode CV6 launch uses the TMA-staged persistent kernel (cp.async.bulk + mbarrier
    ring); B200KF_TMA=0 selects the plain LDG kernel.  Same inputs -> same results, both = oracle.
    N spans 3 CTAs with a partial last one, T is odd (tail chunk)."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 2000, 101
    m, types, dts, payload, x0, P0 = make_case("CV6", N, T, seed=17)
    monkeypatch.setenv("B200KF_SYM", "0")  # per-chain covariances in both kernels: identical arithmetic
    monkeypatch.setenv("B200KF_TMA", "1")
    tma = Engine("CV6", precision=precision).run_batch(types, dts, payload, x0, P0)
    monkeypatch.setenv("B200KF_TMA", "0")
    ldg = Engine("CV6", precision=precision).run_batch(types, dts, payload, x0, P0)
    for k in ("x", "P"):
        np.testing.assert_array_equal(tma[k].cpu().numpy(), ldg[k].cpu().numpy())
    np.testing.assert_allclose(tma["agg"].cpu().numpy(), ldg["agg"].cpu().numpy(), rtol=1e-12, atol=1e-12)
    assert int(tma["status"].sum()) == 0
    ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
    tol = RTOL64 if precision == "f64" else RTOL32
    assert rel(tma["x"].cpu().numpy(), ref["x"]) < tol
    assert rel(tma["P"].cpu().numpy(), ref["P"]) < tol
    a = tma["agg"].cpu().numpy()
    assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < tol
    # odd filter count cannot use 16-byte bulk rows: falls back to the LDG kernel, same answer
    monkeypatch.setenv("B200KF_TMA", "1")
    odd = Engine("CV6", precision=precision).run_batch(types, dts, payload[:, :, :1999], x0[:1999], P0[:1999])
    np.testing.assert_array_equal(odd["P"].cpu().numpy(), tma["P"].cpu().numpy()[:1999])


def test_axis_symmetric_kernel_matches_per_chain_kernel_and_oracle(monkeypatch):
    """CV6 with the reference's axis-symmetric P0: AUTO evaluates the covariance recursion once per
    filter (Sym2Filter).  B200KF_SYM=0 keeps one covariance per chain.  Same numbers either way; a
    P0 that differs between axes must NOT take the symmetric path."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 1800, 157
    m, types, dts, payload, x0, P0 = make_case("CV6", N, T, seed=29)
    ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
    monkeypatch.setenv("B200KF_SYM", "1")
    sym = Engine("CV6").run_batch(types, dts, payload, x0, P0)
    monkeypatch.setenv("B200KF_SYM", "0")
    dec = Engine("CV6").run_batch(types, dts, payload, x0, P0)
    for out in (sym, dec):
        assert int(out["status"].sum()) == 0
        assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL64
        assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL64
        a = out["agg"].cpu().numpy()
        assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL64
    np.testing.assert_allclose(sym["P"].cpu().numpy(), dec["P"].cpu().numpy(), rtol=1e-13, atol=0)
    np.testing.assert_allclose(sym["x"].cpu().numpy(), dec["x"].cpu().numpy(), rtol=1e-12, atol=1e-12)
    # asymmetric axes (one filter is enough): the launch must fall back to per-chain covariances
    monkeypatch.setenv("B200KF_SYM", "1")
    P1 = P0.copy()
    P1[N // 3, 1, 1] *= 1.5
    P1[N // 3, 4, 4] *= 0.5
    ref1 = ko.run_batch(m, types, dts, payload, x0, P1, want_traj=False)
    out1 = Engine("CV6").run_batch(types, dts, payload, x0, P1)
    assert rel(out1["P"].cpu().numpy(), ref1["P"]) < RTOL64
    assert rel(out1["x"].cpu().numpy(), ref1["x"]) < RTOL64
    # different noise per axis in the MODEL disables it on the host side
    from sensorfusion_kalmanfilter_b200.models import default_spec
    spec = default_spec("CV6")
    spec.r_imu = spec.r_imu.copy()
    spec.r_imu[1] = 77.0

    class M2(ko.CV6):
        def R_imu(self):
            return np.diag([50.0, 77.0, 50.0, 10.0, 10.0, 10.0])
    ref2 = ko.run_batch(M2(), types, dts, payload, x0, P0, want_traj=False)
    out2 = Engine(spec).run_batch(types, dts, payload, x0, P0)
    assert rel(out2["P"].cpu().numpy(), ref2["P"]) < RTOL64
    assert rel(out2["x"].cpu().numpy(), ref2["x"]) < RTOL64


@pytest.mark.parametrize("name", ["CV6", "CA15", "BIAS9", "CA8", "CVBU6"])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_shared_covariance_kernel_matches_symmetric_kernel_and_oracle(name, precision, monkeypatch):
    """All filters start from the SAME P (Monte-Carlo seeds of one configuration): AUTO runs the
    covariance / gain recursion once per launch (kf_run_shared_kernel).  B200KF_SHARED=0 keeps one
    covariance per filter.  Same numbers either way, both = oracle; N spans 3 CTAs with a partial
    last one and T is odd.  A launch in which ONE filter's P differs must not take the shared path."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 2000, 151
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=41)
    tol = RTOL64 if precision == "f64" else RTOL32
    ref = ko.run_batch(m, types, dts, payload, x0, P0)
    monkeypatch.setenv("B200KF_SHARED", "1")
    sh = Engine(name, precision=precision).run_batch(types, dts, payload, x0, P0)
    shf = Engine(name, precision=precision).run_batch(types, dts, payload, x0, P0, traj_ncomp=m.n, want_logdet=True,
                                                      mode="shared")
    monkeypatch.setenv("B200KF_SHARED", "0")
    sym = Engine(name, precision=precision).run_batch(types, dts, payload, x0, P0)
    for out in (sh, shf, sym):
        assert int(out["status"].sum()) == 0
        assert rel(out["x"].cpu().numpy(), ref["x"]) < tol
        assert rel(out["P"].cpu().numpy(), ref["P"]) < tol
        a = out["agg"].cpu().numpy()
        assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < tol
        assert np.all(a[:, 3] == T)
    assert rel(shf["traj"].cpu().numpy(), ref["traj"]) < tol
    ld = shf["logdet"].cpu().numpy()
    assert np.max(np.abs(ld - ref["logdet"]) / np.maximum(1.0, np.abs(ref["logdet"]))) < tol
    if precision == "f64" and name == "CV6":  # the two CV6 kernels evaluate the same formulas
        np.testing.assert_allclose(sh["x"].cpu().numpy(), sym["x"].cpu().numpy(), rtol=1e-12, atol=1e-12)
    # one filter with a different (still axis-symmetric) covariance: per-filter recursion for the launch
    monkeypatch.setenv("B200KF_SHARED", "1")
    P1 = P0.copy()
    P1[N - 3] *= 1.5
    ref1 = ko.run_batch(m, types, dts, payload, x0, P1, want_traj=False)
    out1 = Engine(name, precision=precision).run_batch(types, dts, payload, x0, P1)
    assert rel(out1["P"].cpu().numpy(), ref1["P"]) < tol
    assert rel(out1["x"].cpu().numpy(), ref1["x"]) < tol
    a = out1["agg"].cpu().numpy()
    assert np.max(np.abs(a[:, :3] - ref1["agg"]) / np.maximum(1.0, np.abs(ref1["agg"]))) < tol


@pytest.mark.parametrize("name", ["CV6", "CA15", "BIAS9", "CA8", "CVBU6"])
def test_shared_covariance_chunked_continuation_and_predict_only(name):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 900, 240
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=43)
    types = types.copy()
    types[5::11] = 2  # predict-only events
    ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
    eng = Engine(name)
    x, P, agg = x0, P0, None
    for k0 in range(0, T, 100):
        sl = slice(k0, min(T, k0 + 100))
        out = eng.run_batch(types[sl], dts[sl], payload[sl], x, P, agg=agg, mode="shared")
        x, P, agg = out["x"], out["P"], out["agg"]
    assert rel(x.cpu().numpy(), ref["x"]) < RTOL64
    assert rel(P.cpu().numpy(), ref["P"]) < RTOL64
    a = agg.cpu().numpy()
    assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL64
    assert np.all(a[:, 3] == T)


@pytest.mark.parametrize("N", [33150, 40000, 70002])
def test_ring_kernels_at_every_cta_size(N, monkeypatch):
    """The ring kernels pick 224 / 448 / 896 filters per CTA from the shard size (every SM gets a CTA
    when a GPU's shard is small, e.g. BASELINE configs[4] on 8 GPUs = 32768 filters per GPU).  One
    launch of each size class, per-filter covariance and shared covariance, against the oracle."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    T = 45
    for name in ("CV6", "BIAS9", "CVBU6"):
        m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=51)
        ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
        # AUTO keeps shards below the 896-filter CTA regime on the per-filter ring kernel: force the
        # shared-covariance kernel explicitly to cover its small-CTA instantiations too
        for mode, shared in (("shared", "1"), ("auto", "1"), ("auto", "0")):
            monkeypatch.setenv("B200KF_SHARED", shared)
            out = Engine(name).run_batch(types, dts, payload, x0, P0, mode=mode)
            assert int(out["status"].sum()) == 0
            assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL64
            assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL64
            a = out["agg"].cpu().numpy()
            assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL64


@pytest.mark.parametrize("N", [33152, 40000, 70004])
def test_ring_kernels_at_every_cta_size_f32(N, monkeypatch):
    """The float32 instantiations of the same kernels (their rings hold twice the steps per stage, the
    records and the lean aggregate are float32), against the oracle at the float32 tolerance."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    T = 45
    for name in ("CV6", "BIAS9", "CVBU6"):
        m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=53)
        ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
        for mode, shared in (("shared", "1"), ("auto", "0")):
            monkeypatch.setenv("B200KF_SHARED", shared)
            out = Engine(name, precision="f32").run_batch(types, dts, payload, x0, P0, mode=mode)
            assert int(out["status"].sum()) == 0
            assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL32
            assert rel(out["P"].cpu().numpy(), ref["P"]) < RTOL32
            a = out["agg"].cpu().numpy()
            assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL32


@pytest.mark.parametrize("name", ["CA15", "CA8", "CV6"])
@pytest.mark.parametrize("scale", [1e-8, 1e-3, 1e2])  # (beyond ~1e3 the reference's (I - KH) P itself loses digits)
def test_shared_covariance_joint_update_over_covariance_scales(name, scale):
    """The chain warps update a fully measured block jointly (P+ = R S^-1 P through the cofactors of
    S = P + R).  Covariances far below and far above the measurement variances -- where one or the other
    term of S dominates -- against the reference formulation (inv(S)) of the oracle."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    N, T = 64, 120
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=71)
    P0 = P0 * scale
    ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
    out = Engine(name).run_batch(types, dts, payload, x0, P0, mode="shared")
    assert int(out["status"].sum()) == 0
    assert rel(out["x"].cpu().numpy(), ref["x"]) < RTOL64
    P, Pr = out["P"].cpu().numpy(), ref["P"]
    d = np.sqrt(np.abs(np.einsum("fii->fi", Pr)))
    assert np.max(np.abs(P - Pr) / (d[:, :, None] * d[:, None, :])) < RTOL64  # entry by entry, in correlation units
    a = out["agg"].cpu().numpy()
    assert np.max(np.abs(a[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < RTOL64


def test_zero_measurement_variance_stays_exact():
    """A perfect sensor (r = 0 on the IMU velocity pseudo-measurement): the lean symmetric step and the
    shared-covariance kernels divide by the variances, so such a model must take the scalar-update
    kernels (which only divide by P_ss + r) -- and still match the reference formulation (inv(S))."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    from sensorfusion_kalmanfilter_b200.models import ModelSpec, default_spec
    base = default_spec("CV6")
    r_imu = base.r_imu.copy()
    r_imu[3:] = 0.0
    spec = ModelSpec("CV6", "f64", base.q, base.r_gps, r_imu, base.accel_coef)

    class CV6PerfectVelocity(type(ko.get_model("CV6"))):
        def R_imu(self):
            return np.diag(r_imu)

    m = CV6PerfectVelocity()
    N, T = 2000, 60
    _, types, dts, payload, x0, P0 = make_case("CV6", N, T, seed=61)
    ref = ko.run_batch(m, types, dts, payload, x0, P0, want_traj=False)
    for P in (P0, P0 * (1.0 + np.arange(N) / N)[:, None, None]):  # identical and per-filter covariances
        refp = ref if P is P0 else ko.run_batch(m, types, dts, payload, x0, P, want_traj=False)
        out = Engine(spec).run_batch(types, dts, payload, x0, P)
        assert np.isfinite(out["P"].cpu().numpy()).all()
        assert rel(out["x"].cpu().numpy(), refp["x"]) < RTOL64
        # P has exact-zero velocity variance after an IMU event: compare absolutely on the scale of P0
        assert np.max(np.abs(out["P"].cpu().numpy() - refp["P"])) < 1e-9 * np.max(np.abs(refp["P"]) + 1.0)


@pytest.mark.parametrize("name,N", [("CV6", 2000), ("CV6", 33), ("CA15", 2000), ("BIAS9", 900), ("CVBU6", 700)])
def test_step_rows_reads_a_selection_in_place(name, N, monkeypatch):
    """b200kf_run_batch_rows: step k reads payload row step_rows[k] of a longer stream (what the scheduled
    filter does with its selected events).  Same bits as a launch on the gathered copy, on every kernel family
    (ring / LDG, per-filter and shared covariance) and for the precomputed-gain path."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    E, T = 120, 47
    m, types_all, dts_all, payload, x0, P0 = make_case(name, N, E, seed=71)
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(E, T, replace=False)).astype(np.int32)
    types, dts = types_all[rows], dts_all[:T]
    for shared in ("1", "0"):
        monkeypatch.setenv("B200KF_SHARED", shared)
        eng = Engine(name)
        a = eng.run_batch(types, dts, payload[rows], x0, P0, traj_ncomp=m.n)
        b = eng.run_batch(types, dts, payload, x0, P0, traj_ncomp=m.n, step_rows=rows)
        c = eng.run_batch(types, dts, payload, x0, P0, step_rows=rows)  # aggregate-only kernels
        d = eng.run_batch(types, dts, payload[rows], x0, P0)
        for k in ("x", "P", "agg", "traj"):
            np.testing.assert_array_equal(a[k].cpu().numpy(), b[k].cpu().numpy())
        for k in ("x", "P", "agg"):
            np.testing.assert_array_equal(c[k].cpu().numpy(), d[k].cpu().numpy())
    ref = ko.run_batch(m, types, dts, payload[rows], x0, P0, want_traj=False)
    assert rel(b["x"].cpu().numpy(), ref["x"]) < RTOL64


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("N", [600, 33150, 70002, 140000])
def test_output_ring_bulk_stores_match_per_thread_stores_and_oracle(precision, N, monkeypatch):
    """Trajectory mode of the ring kernels: per-step outputs staged in a shared-memory output ring and
    written with cp.async.bulk.global.shared (default) against the per-thread-store form of the same kernel
    (B200KF_BULK_OUT=0) -- bit-identical -- and against the oracle, at every CTA size (224 / 448 / 896 filters,
    ragged last CTA), odd step counts (partial last chunk) and every combination of outputs."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    T = 53
    tol = RTOL64 if precision == "f64" else RTOL32
    for name in ("CV6", "BIAS9", "CVBU6"):
        m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=61)
        pick = np.unique(np.concatenate([np.arange(0, N, max(1, N // 37)), [N - 1]]))
        ref = ko.run_batch(m, types, dts, payload[:, :, pick], x0[pick], P0[pick])
        for ncomp, want_ld in ((6, True), (3, False), (0, True), (m.n if m.n <= 7 else 7, True)):
            outs = {}
            for bulk in ("1", "0"):
                monkeypatch.setenv("B200KF_BULK_OUT", bulk)
                outs[bulk] = Engine(name, precision=precision).run_batch(types, dts, payload, x0, P0, traj_ncomp=ncomp,
                                                                         want_logdet=want_ld, mode="symmetric")
            a, b = outs["1"], outs["0"]
            assert int(a["status"].sum()) == 0
            for k in ("x", "P", "agg") + (("traj",) if ncomp else ()) + (("logdet",) if want_ld else ()):
                assert bool((a[k] == b[k]).all()), (name, ncomp, want_ld, k)
            assert rel(a["x"].cpu().numpy()[pick], ref["x"]) < tol
            if ncomp:
                assert rel(a["traj"].cpu().numpy()[:, :, pick], ref["traj"][:, :ncomp]) < tol
            if want_ld:
                ld = a["logdet"].cpu().numpy()[:, pick]
                assert np.max(np.abs(ld - ref["logdet"]) / np.maximum(1.0, np.abs(ref["logdet"]))) < tol


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", ["CA15", "CA8"])
@pytest.mark.parametrize("N", [1000, 40000])
def test_constant_acceleration_ring_kernel_matches_ldg_kernel_and_oracle(name, N, precision, monkeypatch):
    """The reference's own models (kf_workers.py:493-614 CA15, hw5_2.py:219-304 CA8) with one covariance per
    filter run on the per-filter ring kernel in its constant-acceleration form (SymCAFilter: R-normalised 3 x 3 and
    2 x 2 covariances, joint closed-form H = I update); B200KF_CA_RING=0 keeps the scalar-update LDG kernel.
    Both against the oracle and against each other: aggregate mode, trajectory mode through the output ring,
    time-chunked continuation, predict-only events and the adaptive gate, at both CTA sizes (224 / 448)."""
    from sensorfusion_kalmanfilter_b200.engine import Engine
    T = 61
    tol = RTOL64 if precision == "f64" else RTOL32
    m, types, dts, payload, x0, P0 = make_case(name, N, T, seed=71)
    rng = np.random.default_rng(5)
    P0 = P0 * (1.0 + rng.random(N))[:, None, None]  # one covariance per filter (axis-symmetric, chain-compatible)
    types = types.copy()
    types[[17, 40]] = 2  # predict-only events in the stream
    pick = np.unique(np.concatenate([np.arange(0, N, max(1, N // 41)), [N - 1]]))
    ref = ko.run_batch(m, types, dts, payload[:, :, pick], x0[pick], P0[pick])
    outs = {}
    for ring in ("1", "0"):
        monkeypatch.setenv("B200KF_CA_RING", ring)
        eng = Engine(name, precision=precision)
        agg = eng.run_batch(types, dts, payload, x0, P0)
        full = eng.run_batch(types, dts, payload, x0, P0, traj_ncomp=6, want_logdet=True)
        x, P, a = x0, P0, None
        for k0 in range(0, T, 25):
            sl = slice(k0, min(T, k0 + 25))
            o = eng.run_batch(types[sl], dts[sl], payload[sl], x, P, agg=a)
            x, P, a = o["x"], o["P"], o["agg"]
        outs[ring] = (agg, full, o)
        for out in (agg, full, o):
            assert int(out["status"].sum()) == 0
            assert rel(out["x"].cpu().numpy()[pick], ref["x"]) < tol
            assert rel(out["P"].cpu().numpy()[pick], ref["P"]) < tol
            ag = out["agg"].cpu().numpy()[pick]
            assert np.max(np.abs(ag[:, :3] - ref["agg"]) / np.maximum(1.0, np.abs(ref["agg"]))) < tol
            assert np.all(ag[:, 3] == T)
        assert rel(full["traj"].cpu().numpy()[:, :, pick], ref["traj"][:, :6]) < tol
        ld = full["logdet"].cpu().numpy()[:, pick]
        assert np.max(np.abs(ld - ref["logdet"]) / np.maximum(1.0, np.abs(ref["logdet"]))) < tol
    # the two kernels agree far below the oracle tolerance (same recursion, different formulation)
    t2 = 1e-11 if precision == "f64" else 2e-5
    for a, b in zip(outs["1"], outs["0"]):
        assert rel(a["x"].cpu().numpy(), b["x"].cpu().numpy()) < t2
        assert rel(a["P"].cpu().numpy(), b["P"].cpu().numpy()) < t2
    if precision == "f64" and N <= 1000:
        # adaptive gate on the ring: the same decisions as the LDG kernel away from the threshold
        thr = float(np.median(ref["logdet"][20:])) + 1.0
        g = {}
        for ring in ("1", "0"):
            monkeypatch.setenv("B200KF_CA_RING", ring)
            g[ring] = Engine(name).run_batch(types, dts, payload, x0, P0, gate_threshold=thr, want_used=True, want_logdet=True)
        ua, ub = g["1"]["used"].cpu().numpy(), g["0"]["used"].cpu().numpy()
        same = (ua == ub).all(axis=0)
        assert same.mean() > 0.99 and 0.02 < ua.mean() < 0.98
        assert rel(g["1"]["x"].cpu().numpy()[same], g["0"]["x"].cpu().numpy()[same]) < 1e-10
        assert rel(g["1"]["P"].cpu().numpy()[same], g["0"]["P"].cpu().numpy()[same]) < 1e-10
