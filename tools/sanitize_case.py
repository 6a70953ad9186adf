"""Small invocations of every ring / search kernel for `compute-sanitizer --tool memcheck` (one tool per
gpurun call).  python tools/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import synth  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402
from sensorfusion_kalmanfilter_b200.models import default_spec  # noqa: E402

rng = np.random.default_rng(0)
for name in ("CV6", "BIAS9", "CA8", "CA15"):
    sp = default_spec(name)
    N, T = 1000, 37
    types, dts = synth.gps_imu_schedule(T, gps_every=5)
    pay = synth.make_payload(sp.d, types, N, seed=1, m_gps=sp.m_gps)
    x0 = rng.normal(0, 1, (N, sp.n))
    P0 = np.repeat(sp.default_P0()[None], N, axis=0)
    P1 = P0 * (1.0 + np.arange(N) / N)[:, None, None]
    for prec in ("f64", "f32"):
        eng = Engine(name, precision=prec)
        a = eng.run_batch(types, dts, pay, x0, P0)                                    # shared covariance
        b = eng.run_batch(types, dts, pay, x0, P1)                                    # per-filter covariance
        c = eng.run_batch(types, dts, pay, x0, P0, traj_ncomp=sp.n, want_logdet=True)  # per-step outputs
        print(name, prec, float(a["agg"][:, 1].mean()), float(b["agg"][:, 1].mean()), float(c["logdet"].mean()))
e = Engine("CA15")
L = 15
st = np.where(np.arange(L) % 4 == 1, 0, 1).astype(np.uint8)
tm = 10.0 + np.cumsum(0.005 + 0.001 * rng.random(L))
best, _ = e.score_schedules(st, tm, 9.99, float(tm[-1]), e.spec.default_P0() * 1e-2, first=3, count=(1 << L) - 7,
                            threshold=-60.0, want_per_schedule=True)
print("search", best["argmin_sum_index"], best["n_feasible"])
q = [3, 2, 4, 3]
E = sum(q)
b2, _ = e.score_products(q, (rng.random(E) < 0.8).astype(np.uint8), 5.0 + np.sort(rng.uniform(0, 0.1, E)),
                         rng.normal(0, 0.1, (E, 9)), rng.normal(0, 1, (E, 3)), np.zeros(15), e.spec.default_P0(),
                         want_per_combo=True)
print("products", b2["index"])
# round 2: the three-lanes-per-filter dense kernel (ragged filter counts, all outputs, mixed gate decisions),
# the constant-acceleration ring kernels and the warp-per-instance selection with deferred gain records
e = Engine("CA15")
for N in (1, 13, 250):
    T = 23
    types, dts = synth.gps_imu_schedule(T, gps_every=5)
    pay = synth.make_payload(9, types, N, seed=2, m_gps=3)
    A = rng.normal(size=(N, 15, 15))
    Pd = A @ np.swapaxes(A, 1, 2) + 5.0 * np.eye(15)
    x0 = rng.normal(0, 1, (N, 15))
    a = e.run_batch(types, dts, pay, x0, Pd)
    b = e.run_batch(types, dts, pay, x0, Pd, traj_ncomp=15, want_logdet=True, want_ptraj=True, want_used=True)
    thr = np.linspace(float(b["logdet"].min()), float(b["logdet"].max()), N)
    c = e.run_batch(types, dts, pay, x0, Pd, gate_threshold=thr, want_used=True)
    print("dense", N, float(a["agg"][:, 1].mean()), float(b["logdet"].mean()), float(c["used"].float().mean()))
ev_t, ev_dt = synth.gps_imu_schedule(700, gps_every=9)
times = 50.0 + np.cumsum(np.maximum(ev_dt, 1e-4))
rates = [10.0, 45.0, 120.0]
P0 = np.repeat(e.spec.default_P0()[None], len(rates), axis=0)
sels, Pf, recs = e.select_greedy(ev_t, times, [1.0 / r for r in rates], [49.99] * len(rates), P0, want_gains=True)
st = e.select_state(ev_t, times, [1.0 / r for r in rates], [49.99] * len(rates), P0)
for e_end in (100, 333, 700):
    st.advance(e_end)
print("select", [len(s) for s in sels], float(Pf.sum()), None if recs is None else int(recs.shape[1]))
