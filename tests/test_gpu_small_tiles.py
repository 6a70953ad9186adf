"""b200kf_run_gains with B200KF_RUN_SMALL_TILES (128 filters per CTA): same numbers as the default tiling, and both
equal the launch that runs the covariance recursion itself (checked against the oracle elsewhere)."""
import numpy as np
import pytest

from sensorfusion_kalmanfilter_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,precision", [("CA15", "f64"), ("CV6", "f64"), ("CA8", "f32")])
def test_small_tiles_are_bit_identical(name, precision):
    from sensorfusion_kalmanfilter_b200.engine import Engine
    eng = Engine(name, precision=precision)
    sp = eng.spec
    N = 1000 if precision == "f64" else 1004  # ragged last CTA for both tilings (16-byte rows)
    ev_t, ev_dt = synth.gps_imu_schedule(900, gps_every=9)
    times = 50.0 + np.cumsum(np.maximum(ev_dt, 1e-4))
    sels, Pf, recs = eng.select_greedy(ev_t, times, [1.0 / 60.0], [49.99], sp.default_P0()[None], want_gains=True)
    assert recs is not None
    sel = sels[0]
    T = len(sel)
    pay = synth.make_payload(sp.d, ev_t, N, seed=3, m_gps=sp.m_gps)
    x0 = np.random.default_rng(1).normal(0, 1, (N, sp.n))
    outs = []
    for small in (False, True):
        outs.append(eng.run_gains(recs[0, :T], pay, x0, traj_ncomp=min(6, sp.n), want_logdet=True,
                                  step_rows=sel.astype(np.int32), small_tiles=small))
    a, b = outs
    for k in ("x", "agg", "traj", "logdet", "status"):
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy()), k
    # and against the recursion-in-the-launch path (shared covariance: every filter starts from the same P0)
    dts = np.diff(times[sel], prepend=49.99)
    P0 = np.repeat(sp.default_P0()[None], N, axis=0)
    c = eng.run_batch(ev_t[sel], dts, pay, x0, P0, step_rows=sel.astype(np.int32))
    tol = 1e-9 if precision == "f64" else 1e-4
    xa, xc = a["x"].cpu().numpy().astype(np.float64), c["x"].cpu().numpy().astype(np.float64)
    assert np.max(np.abs(xa - xc)) / np.max(np.abs(xc)) < tol
