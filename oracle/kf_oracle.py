"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched GPS/IMU Kalman-filter hot path.

A numpy-float64 restatement of the reference's algorithm (IseanB/SensorFusion-KalmanFilter),
written to be *checked against*, never shipped: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product package
(sensorfusion-kalmanfilter_b200/) must not.

Parity pin: tests/test_oracle_golden.py checks every function here against
tests/golden/*.npz, which oracle/make_golden.py produced by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_import.py) on seeded synthetic streams.
The reference itself ships no tests or golden vectors for this path, and its
saved notebook outputs need the absent imu_data.csv, so "outputs of the reference itself run
here" is the pin.  The CV6, BIAS9 and CVBU6 models do not exist in the reference (SURVEY.md 8(d)
defines them); for those the oracle is the definition and parity is "restatement only".

Everything works on a leading batch axis: x is [..., n], P is [..., n, n]; a single filter is
a batch of one.  The four formulas are kept in the reference's own (non-symmetrised,
inverse-based) form:

  predict   x = F x ; P = F P F^T + Q                          kf_workers.py:546-549, 690-691
  gain      K = (P H^T) inv(H P H^T + R)                       kf_workers.py:616-621
  update    y = z - H x ; x = x + K y ; P = (I - K H) P        kf_workers.py:709-711
  metric    sign, logabsdet = slogdet(P)                       kf_workers.py:716 / :70-71
"""
from __future__ import annotations

import itertools
import math

import numpy as np

GPS, IMU, PREDICT_ONLY = 0, 1, 2  # event type codes used by the packed (array) form


# --------------------------------------------------------------------------------------
# Models
# --------------------------------------------------------------------------------------
class OracleModel:
    """One linear GPS/IMU model: F(dt), Q(dt), H/R per sensor and the IMU pseudo-measurement."""

    name = "?"
    n = 0          # state dimension
    m_gps = 0      # GPS measurement dimension
    d = 0          # payload doubles per event in the packed form
    imu_is_control = False  # True: IMU event is x = F x + B u with no update (BIAS9)

    def F(self, dt):
        raise NotImplementedError

    def Q(self, dt):
        raise NotImplementedError

    def H_gps(self):
        H = np.zeros((self.m_gps, self.n))
        H[np.arange(self.m_gps), np.arange(self.m_gps)] = 1
        return H

    def R_gps(self):
        return np.diag([3.0] * self.m_gps)

    def H_imu(self):
        return np.eye(self.n)

    def R_imu(self):
        raise NotImplementedError

    def P0(self):
        raise NotImplementedError

    def imu_z(self, x, u, dt):
        """Pseudo-measurement vector built from the PREDICTED state x [..., n] and payload u [..., d]."""
        raise NotImplementedError

    def B(self, dt):
        raise NotImplementedError


class CA15(OracleModel):
    """15-state 3-D model of kf_workers.py: [p(3), rpy(3), v(3), w(3), a(3)]."""

    name, n, m_gps, d = "CA15", 15, 3, 9

    def F(self, dt):  # kf_workers.py:493-517
        F = np.eye(15)
        for ax in range(3):
            F[ax, 6 + ax] = dt
            F[ax, 12 + ax] = 0.5 * dt ** 2
            F[3 + ax, 9 + ax] = dt
            F[6 + ax, 12 + ax] = dt
        return F

    def Q(self, dt):  # kf_workers.py:519-544
        return np.diag([5 * dt] * 3 + [0.05 * dt] * 3 + [1 * dt] * 3 + [0.1 * dt] * 3 + [2 * dt] * 3)

    def R_imu(self):  # kf_workers.py:587-614
        return np.diag([50.0] * 3 + [0.05] * 3 + [10.0] * 3 + [0.1] * 3 + [100.0] * 3)

    def P0(self):  # kf_workers.py:651
        return np.diag([10000.0] * 3 + [1000.0] * 3 + [1000.0] * 3 + [1000.0] * 3 + [10000.0] * 3)

    def imu_z(self, x, u, dt):  # kf_workers.py:698-704; payload u = [r,p,y, wx,wy,wz, ax,ay,az]
        acc = u[..., 6:9]
        V = x[..., 6:9] + acc * dt
        X = x[..., 0:3] + V * dt
        return np.concatenate([X, u[..., 0:3], V, u[..., 3:6], acc], axis=-1)


class CA8(OracleModel):
    """8-state 2-D model of hw5_2.py: [x, y, th, vx, vy, thdot, ax, ay]."""

    name, n, m_gps, d = "CA8", 8, 2, 4

    def F(self, dt):  # hw5_2.py:219-231
        F = np.eye(8)
        F[0, 3] = F[1, 4] = F[2, 5] = F[3, 6] = F[4, 7] = dt
        F[0, 6] = F[1, 7] = 0.5 * dt ** 2
        return F

    def Q(self, dt):  # hw5_2.py:233-251
        return np.diag([5 * dt, 5 * dt, 0.05 * dt, 1 * dt, 1 * dt, 0.1 * dt, 2 * dt, 2 * dt])

    def R_imu(self):  # hw5_2.py:286-304
        return np.diag([50.0, 50.0, 0.05, 10.0, 10.0, 0.1, 100.0, 100.0])

    def P0(self):  # hw5_2.py:317-326
        return np.diag([1000.0, 1000.0, 100.0, 100.0, 100.0, 100.0, 1000.0, 1000.0])

    def imu_z(self, x, u, dt):  # hw5_2.py:362-370; payload u = [yaw, wz, ax, ay]
        acc = u[..., 2:4]
        V = x[..., 3:5] + acc * dt
        X = x[..., 0:2] + V * dt
        return np.concatenate([X, u[..., 0:1], V, u[..., 1:2], acc], axis=-1)


class CV6(OracleModel):
    """6-state constant-velocity model [p(3), v(3)] = CA15 restricted to rows/cols {0,1,2,6,7,8}
    (formulas of kf_workers.py:501-509, 521-523, 583, 589-590, 651)."""

    name, n, m_gps, d = "CV6", 6, 3, 3

    def F(self, dt):
        F = np.eye(6)
        F[0, 3] = F[1, 4] = F[2, 5] = dt
        return F

    def Q(self, dt):
        return np.diag([5 * dt] * 3 + [1 * dt] * 3)

    def R_imu(self):
        return np.diag([50.0] * 3 + [10.0] * 3)

    def P0(self):
        return np.diag([10000.0] * 3 + [1000.0] * 3)

    def imu_z(self, x, u, dt):  # payload u = [ax, ay, az]
        V = x[..., 3:6] + u[..., 0:3] * dt
        X = x[..., 0:3] + V * dt
        return np.concatenate([X, V], axis=-1)


class BIAS9(OracleModel):
    """9-state [p(3), v(3), b(3)] accelerometer-bias model: IMU
    acceleration is a control input (x = F x + B u, no update); GPS events predict with u = 0
    and then update the position rows."""

    name, n, m_gps, d = "BIAS9", 9, 3, 3
    imu_is_control = True

    def F(self, dt):
        F = np.eye(9)
        for ax in range(3):
            F[ax, 3 + ax] = dt
            F[ax, 6 + ax] = -0.5 * dt ** 2
            F[3 + ax, 6 + ax] = -dt
        return F

    def B(self, dt):
        B = np.zeros((9, 3))
        for ax in range(3):
            B[ax, ax] = 0.5 * dt ** 2
            B[3 + ax, ax] = dt
        return B

    def Q(self, dt):
        return np.diag([5 * dt] * 3 + [1 * dt] * 3 + [2 * dt] * 3)

    def R_imu(self):
        return np.zeros((9, 9))

    def P0(self):
        return np.diag([10000.0] * 3 + [1000.0] * 3 + [10000.0] * 3)


class CVBU6(OracleModel):
    """6-state [p(3), v(3)] model in the north star's explicit-control form x = F x + B u, P = F P F^T + Q
    ("explicit-Bu variant"): B = [[dt^2/2 I], [dt I]], IMU acceleration is the control
    input (no IMU update), GPS events predict with u = 0 and update the position rows.  F, Q, P0 as CV6."""

    name, n, m_gps, d = "CVBU6", 6, 3, 3
    imu_is_control = True

    def F(self, dt):
        F = np.eye(6)
        F[0, 3] = F[1, 4] = F[2, 5] = dt
        return F

    def B(self, dt):
        B = np.zeros((6, 3))
        for ax in range(3):
            B[ax, ax] = 0.5 * dt ** 2
            B[3 + ax, ax] = dt
        return B

    def Q(self, dt):
        return np.diag([5 * dt] * 3 + [1 * dt] * 3)

    def R_imu(self):
        return np.zeros((6, 6))

    def P0(self):
        return np.diag([10000.0] * 3 + [1000.0] * 3)


MODELS = {"CA15": CA15, "CA8": CA8, "CV6": CV6, "BIAS9": BIAS9, "CVBU6": CVBU6}


def get_model(name):
    return MODELS[name]()


# --------------------------------------------------------------------------------------
# The four formulas (batched over leading axes)
# --------------------------------------------------------------------------------------
def predict(model, x, P, dt, u=None):
    """kf_workers.py:688-691 / :546-549.  With u (BIAS9 IMU event) adds B u."""
    F = model.F(dt)
    Q = model.Q(dt)
    x = np.matmul(x, F.T)
    if u is not None:
        x = x + np.matmul(u, model.B(dt).T)
    P = np.matmul(np.matmul(F, P), F.T) + Q
    return x, P


def kalman_gain(P, H, R):
    """kf_workers.py:616-621 -- explicit inverse, as the reference."""
    S = np.matmul(np.matmul(H, P), H.T) + R
    return np.matmul(np.matmul(P, H.T), np.linalg.inv(S))


def update(x, P, z, H, R):
    """kf_workers.py:708-711 -- non-Joseph, not re-symmetrised."""
    n = P.shape[-1]
    K = kalman_gain(P, H, R)
    y = z - np.matmul(x, H.T)
    x = x + np.matmul(K, y[..., None])[..., 0]
    P = np.matmul(np.eye(n) - np.matmul(K, H), P)
    return x, P


def logdet(P):
    """(sign, logabsdet) of kf_workers.py:716."""
    return np.linalg.slogdet(P)


def step(model, x, P, etype, dt, payload, do_update=True):
    """One filter step = one sensor event (kf_workers.py:681-711)."""
    if model.imu_is_control and etype == IMU:
        return predict(model, x, P, dt, u=payload[..., : model.d])
    x, P = predict(model, x, P, dt)
    if not do_update or etype == PREDICT_ONLY:  # no-update events (kf_workers.py:74-82, :1060-1159)
        return x, P
    if etype == GPS:
        return update(x, P, payload[..., : model.m_gps], model.H_gps(), model.R_gps())
    return update(x, P, model.imu_z(x, payload, dt), model.H_imu(), model.R_imu())


# --------------------------------------------------------------------------------------
# Packed (array-level) batch run: what the C ABI's b200kf_run_batch computes
# --------------------------------------------------------------------------------------
def run_batch(model, types, dts, payload, x0, P0, gate_threshold=None, want_traj=True):
    """Run N filters over a shared schedule.

    types [T] (0 GPS / 1 IMU), dts [T] float64, payload [T, d, N] (time-major, filter fastest),
    x0 [N, n], P0 [N, n, n].  Returns dict with final x, P, per-step traj [T, n_out, N],
    per-step logdet [T, N] and aggregates [N, 3] = (max, sum, last) of logabsdet over the T
    steps.  gate_threshold: adaptive gate of kf_workers.py:1023-1025 (update only when
    sign*logabsdet of the predicted P exceeds it); used [T, N] records the decisions.
    """
    T = len(types)
    x = np.array(x0, dtype=np.float64)
    P = np.array(P0, dtype=np.float64)
    N = x.shape[0]
    lds = np.zeros((T, N))
    used = np.ones((T, N), dtype=bool)
    traj = np.zeros((T, model.n, N)) if want_traj else None
    for k in range(T):
        u = np.ascontiguousarray(payload[k].T)  # [N, d]
        if gate_threshold is None:
            x, P = step(model, x, P, int(types[k]), float(dts[k]), u)
        else:
            xp, Pp = predict(model, x, P, float(dts[k]))
            s, l = logdet(Pp)
            g = (s * l) > gate_threshold
            xu, Pu = step(model, x, P, int(types[k]), float(dts[k]), u)
            x = np.where(g[:, None], xu, xp)
            P = np.where(g[:, None, None], Pu, Pp)
            used[k] = g
        lds[k] = logdet(P)[1]
        if want_traj:
            traj[k] = x.T
    agg = np.stack([lds.max(axis=0), lds.sum(axis=0), lds[-1]], axis=1) if T else np.zeros((N, 3))
    return {"x": x, "P": P, "traj": traj, "logdet": lds, "agg": agg, "used": used}


# --------------------------------------------------------------------------------------
# Event-level loops over the reference's `indexed_sensor_data` layout
#   (i, 'GPS', t, {'easting','northing','altitude',...}) kf_workers.py:331
#   (i, 'IMU', t, [str t, r,p,y, wx,wy,wz, ax,ay,az])     kf_workers.py:367
# --------------------------------------------------------------------------------------
def event_payload(model, stype, sdata):
    """Extract the packed payload of one event for `model` (what the host packer ships)."""
    u = np.zeros(model.d)
    if stype == "GPS":
        vals = [sdata["easting"], sdata["northing"], sdata.get("altitude", 0.0)][: model.m_gps]
        u[: model.m_gps] = vals
    elif model.name == "CA15":
        u[:] = [float(v) for v in sdata[1:10]]
    elif model.name == "CA8":
        u[:] = [sdata[3], sdata[6], sdata[7], sdata[8]]
    else:  # CV6 / BIAS9: acceleration only
        u[:] = [sdata[7], sdata[8], sdata[9]]
    return u


def _type_code(stype):
    return GPS if stype == "GPS" else IMU


def _init_from_first_gps(model, events, start_idx, stop):
    """First-GPS initialisation shared by the run_* loops (kf_workers.py:651-666)."""
    x = np.zeros(model.n)
    for i, (_, stype, t, sdata) in enumerate(events[start_idx:stop]):
        if stype == "GPS":
            x[: model.m_gps] = [sdata["easting"], sdata["northing"], sdata["altitude"]][: model.m_gps]
            return x, t, start_idx + i
    return None, None, -1


def run_full(model, events, start_idx=None, end_idx=None, initial_pt=None, initial_state=None):
    """run_kalman_filter_full, kf_workers.py:623-728 -> (states, logdets, P, prev_time)."""
    if start_idx is None or start_idx < 0:
        start_idx = 0
    if end_idx is None or end_idx > len(events):
        end_idx = len(events)
    if initial_pt is not None and initial_state is not None:
        P = np.array(initial_pt, dtype=np.float64)
        x = np.zeros(model.n)
        x[0:6] = initial_state[1:7]
        prev = initial_state[0]
        first = start_idx
    else:
        P = model.P0()
        x, prev, first = _init_from_first_gps(model, events, start_idx, end_idx + 1)
        if first == -1:
            return [], [], []
    states = [(prev, *x[:6])]
    lds = [logdet(P)[1]]
    for (_, stype, t, sdata) in events[first:end_idx]:
        dt = t - prev
        if dt < 0:
            prev = t
            continue
        x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        states.append((t, *x[:6]))
        lds.append(logdet(P)[1])
        prev = t
    return states, lds, P, prev


def run_legacy(model, events, start_idx, end_idx):
    """run_kalman_filter, kf_workers.py:738-824 -> (states, covs); zero init, starts at the
    first GPS event which is itself processed with dt = 0."""
    x = np.zeros(model.n)
    P = model.P0()
    states = [(0, *x[:6])]
    covs = [P.copy()]
    prev = None
    started = False
    for (_, stype, t, sdata) in events[start_idx:end_idx]:
        if stype == "GPS" and not started:
            started = True
            prev = t
        if not started:
            continue
        dt = t - prev
        x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        states.append((t, *x[:6]))
        covs.append(P.copy())
        prev = t
    return states, covs


def run_adaptive(model, events, start_idx=None, end_idx=None, R_threshold=None,
                 initial_pt=None, initial_state=None, never_update=False):
    """run_adaptive_threshold_kalman_filter, kf_workers.py:959-1058, and (never_update=True)
    run_no_update_kalman_filter, :1060-1159.  Returns
    (states, logdets, P, previous_time, measurement_times).  Note the reference's dt<0 branch
    assigns a dead variable (`prev_time`, :1013-1015) so previous_time is NOT advanced there."""
    if start_idx is None or start_idx < 0:
        start_idx = 0
    if end_idx is None or end_idx > len(events):
        end_idx = len(events)
    if R_threshold is None:
        R_threshold = -float("inf")
    mtimes = []
    if initial_pt is not None and initial_state is not None:
        P = np.array(initial_pt, dtype=np.float64)
        x = np.zeros(model.n)
        x[0:6] = initial_state[1:7]
        prev = initial_state[0]
        first = start_idx
    else:
        P = model.P0()
        x, prev, first = _init_from_first_gps(model, events, start_idx, end_idx)
        if first == -1:
            return None
        mtimes.append(prev)
    states = [(prev, *x[:6])]
    lds = [logdet(P)[1]]
    for (_, stype, t, sdata) in events[first:end_idx]:
        dt = t - prev
        if dt < 0:
            continue
        xp, Pp = predict(model, x, P, dt)
        s, l = logdet(Pp)
        if (not never_update) and (l * s > R_threshold):
            mtimes.append(t)
            x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        else:
            x, P = xp, Pp
        states.append((t, *x[:6]))
        lds.append(logdet(P)[1])
        prev = t
    return states, lds, P, prev, mtimes


# ---- Scheduler (kf_workers.py:99-233) -------------------------------------------------
def cov_matrix(S_rows, Sigma_prev, R, H):
    """Scheduler.cov_matrix cpu branch, kf_workers.py:112-147: posterior covariance after using
    the 1-based sensor rows S_rows of (H, R)."""
    k = len(S_rows)
    if k == R.shape[0]:
        R_hat, H_hat = R, H
    else:
        rows = np.array(sorted(S_rows)) - 1
        R_hat = R[np.ix_(rows, rows)]
        H_hat = H[rows, :]
    S = R_hat + np.dot(H_hat, np.dot(Sigma_prev, H_hat.T))
    K = np.dot(np.dot(Sigma_prev, H_hat.T), np.linalg.inv(S))
    return Sigma_prev - np.dot(np.dot(K, H_hat), Sigma_prev)


def gain(model, stype, Sigma):
    """Scheduler.gain, kf_workers.py:174-185: trace of the posterior using ONLY row 1."""
    if stype == "GPS":
        return np.trace(cov_matrix([1], Sigma, model.R_gps(), model.H_gps()))
    return np.trace(cov_matrix([1], Sigma, model.R_imu(), model.H_imu()))


def greedy_schedule(model, queue, Sigma):
    """Scheduler.greedy_schedule, kf_workers.py:195-213: index of the LARGEST posterior trace,
    first wins ties."""
    best, best_i = -np.inf, None
    for i, ev in enumerate(queue):
        g = gain(model, ev[1], Sigma)
        if g > best:
            best, best_i = g, i
    return best_i


def run_scheduled(model, events, processing_frequency, start_idx=None, end_idx=None,
                  initial_pt=None, initial_state=None, selection="greedy", rng_choice=None):
    """run_kalman_filter_scheduled, kf_workers.py:826-957 -> (states, logdets, P).
    selection='random' draws through rng_choice(n) (the reference calls np.random.choice(n))."""
    if start_idx is None or start_idx < 0:
        start_idx = 0
    if end_idx is None or end_idx > len(events):
        end_idx = len(events)
    if initial_state is not None:
        P = np.array(initial_pt, dtype=np.float64)
        x = np.zeros(model.n)
        x[0:6] = initial_state[1:7]
        first = start_idx
        prev = initial_state[0]
    else:
        P = model.P0()
        x, prev, first = _init_from_first_gps(model, events, start_idx, end_idx)
        if first == -1:
            return None, None
    states = [(prev, *x[:6])]
    lds = [logdet(P)[1]]
    queue = []
    if end_idx == -1:
        end_idx = len(events)
    for ev in events[first + 1:end_idx]:
        t = ev[2]
        if t - prev < 1 / processing_frequency:
            queue.append(ev)
            continue
        if not queue:
            queue.append(ev)
        if selection == "random":
            sel = rng_choice(len(queue))
        else:
            sel = greedy_schedule(model, queue, P)
        (_, stype, st, sdata) = queue[sel]
        queue = []
        dt = st - prev
        x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        states.append((st, *x[:6]))
        lds.append(logdet(P)[1])
        prev = st
    return states, lds, P


# ---- brute-force schedule search (kf_workers.py:22-97, 1218-1392) ---------------------
def worker_combo(model, combo, xt, Pt, prev_time, target_end_time):
    """One combo of evaluate_combo_chunk_worker, kf_workers.py:29-86.
    Returns (traj, x_final, P_final, log_det list)."""
    x = np.array(xt, dtype=np.float64).copy()
    P = np.array(Pt, dtype=np.float64).copy()
    traj = [(prev_time, *x[:6])]
    s, l = logdet(P)
    lds = [s * l]
    cur = prev_time
    for (_, stype, t, sdata) in combo:
        dt = t - cur
        if dt < 0:
            continue
        x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        traj.append((t, *x[:6]))
        cur = t
        s, l = logdet(P)
        lds.append(s * l)
    if cur < target_end_time - 1e-8:
        dt = target_end_time - cur
        x, P = predict(model, x, P, dt)
        traj.append((target_end_time, *x[:6]))
        s, l = logdet(P)
        lds.append(s * l)
    return traj, x, P, lds


def worker_chunk(model, chunk, xt, Pt, prev_time, target_end_time):
    """evaluate_combo_chunk_worker's return list, kf_workers.py:86."""
    out = []
    for combo in chunk:
        traj, x, _, lds = worker_combo(model, combo, xt, Pt, prev_time, target_end_time)
        out.append((0, traj, combo, x.copy(), None, lds, len(combo)))
    return out


def enumerate_subsets(n):
    """Index tuples in the driver's order: cardinality ascending, then
    itertools.combinations lexicographic (kf_workers.py:1325-1336)."""
    for k in range(1, n + 1):
        yield from itertools.combinations(range(n), k)


def subset_mask(idx_tuple):
    m = 0
    for i in idx_tuple:
        m |= 1 << i
    return m


def schedule_trace(model, types, times, mask, P0, t_prev, t_end):
    """Covariance-only log-det trace of one subset schedule (bit i of mask = use event i):
    exactly the log_det list worker_combo returns (it never depends on measurement values,
    kf_workers.py:44,64,67,70)."""
    P = np.array(P0, dtype=np.float64).copy()
    n = model.n
    s, l = logdet(P)
    lds = [s * l]
    cur = t_prev
    for i, (ty, t) in enumerate(zip(types, times)):
        if not (mask >> i) & 1:
            continue
        dt = t - cur
        if dt < 0:
            continue
        _, P = step(model, np.zeros(n), P, int(ty), dt, np.zeros(model.d))
        cur = t
        s, l = logdet(P)
        lds.append(s * l)
    if cur < t_end - 1e-8:
        _, P = predict(model, np.zeros(n), P, t_end - cur)
        s, l = logdet(P)
        lds.append(s * l)
    return lds


def brute_force_select(model, window, xt, Pt, prev_time, target_end_time, R_threshold,
                       chunk_size=500):
    """Selection rule of run_brute_force_kalman_filter_no_sampling_min_usage,
    kf_workers.py:1325-1371: first chunk (of `chunk_size` combos in enumeration order) holding
    a combo with max(log_det) < R_threshold; inside it the first such combo
    (min(..., key=metric) with metric == 0 everywhere keeps the first).  Returns the result
    dict of :1363-1371 plus 'selected_indices', or None (:1391-1392)."""
    n = len(window)
    for k in range(1, n + 1):
        it = itertools.combinations(range(n), k)
        while True:
            idx_chunk = list(itertools.islice(it, chunk_size))
            if not idx_chunk:
                break
            for idx in idx_chunk:
                combo = tuple(window[i] for i in idx)
                traj, x, _, lds = worker_combo(model, combo, xt, Pt, prev_time, target_end_time)
                if max(lds) < R_threshold:
                    return {"selected_sensors": combo, "final_state": x, "final_covariance": None,
                            "trajectory": traj, "accuracy_metric": 0, "log_determinants": lds,
                            "num_measurements_used": len(combo), "selected_indices": idx}
    return None


def lex_rank_key(mask, L):
    """Sort key reproducing the reference enumeration order on bitmasks: (popcount, order
    inside the cardinality class).  Within one cardinality the lexicographically smaller index
    tuple has the LARGER bit-reversed value, hence the complement."""
    rev = 0
    for i in range(L):
        if (mask >> i) & 1:
            rev |= 1 << (L - 1 - i)
    return (bin(mask).count("1"), ((1 << L) - 1) - rev)


# ---- hw5_2.py loops -------------------------------------------------------------------
def hw5_2_run(model, events):
    """hw5_2.KF_SensorFusion.run_kalman_filter, hw5_2.py:313-380 -> [(x, y, theta)]."""
    x = np.zeros(model.n)
    P = model.P0()
    out = [(x[0], x[1], x[2])]
    started = False
    prev = None
    for (_, stype, t, sdata) in events:
        if stype == "GPS" and not started:
            started = True
            prev = t
        if not started:
            continue
        dt = t - prev
        x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        out.append((x[0], x[1], x[2]))
        prev = t
    return out


def hw5_2_dead_reckoning(model, events):
    """hw5_2.KF_SensorFusion.run_dead_reckoning_for_IMU, hw5_2.py:382-436 (IMU events only)."""
    x = np.zeros(model.n)
    P = model.P0()
    out = []
    prev = None
    for (_, stype, t, sdata) in events:
        if stype != "IMU":
            continue
        dt = t - prev if prev is not None else 0
        x, P = step(model, x, P, IMU, dt, event_payload(model, stype, sdata))
        out.append((x[0], x[1], x[2]))
        prev = t
    return out


def comb(n, k):
    return math.comb(n, k)


# ---- notebook brute force: product of per-window queues, RMSE objective -----------------
# (KF_SensorFusion.ipynb:766-802 metric, :804-954 search; raw JSON line numbers)
def nb_P0(n=15):
    """Initial covariance of the notebook's brute force (KF_SensorFusion.ipynb:813)."""
    return np.diag([1000.0] * 3 + [100.0] * 3 + [100.0] * 3 + [100.0] * 3 + [1000.0] * 3)


def bucket_queues(events, start_idx, end_idx, processing_frequency):
    """Step 1 (KF_SensorFusion.ipynb:826-847): one queue per sampling period, starting at the
    first GPS event; the event that closes a non-empty window is dropped and previous_time moves
    to ITS time; a trailing partial queue is discarded."""
    if end_idx is None:
        end_idx = len(events)
    queues, queue = [], []
    started, prev = False, None
    for ev in events[start_idx:end_idx]:
        if not started and ev[1] == "GPS":
            started, prev = True, ev[2]
        if not started:
            continue
        if ev[2] - prev < 1 / processing_frequency:
            queue.append(ev)
            continue
        if len(queue) == 0:
            queue.append(ev)
        queues.append(list(queue))
        queue = []
        prev = ev[2]
    return queues


def lerp_extrapolate(x, xp, fp):
    """scipy interp1d(kind='linear', fill_value='extrapolate') on sorted xp, in the form scipy 1.18
    evaluates it (interp1d._call_linear: side='left' search clipped to [1, n-1], two-weight blend)
    so it is bit-identical to what the reference computes in this container."""
    x, xp, fp = np.asarray(x, float), np.asarray(xp, float), np.asarray(fp, float)
    hi = np.clip(np.searchsorted(xp, x), 1, len(xp) - 1)
    lo = hi - 1
    return ((x - xp[lo]) / (xp[hi] - xp[lo])) * fp[hi] + ((xp[hi] - x) / (xp[hi] - xp[lo])) * fp[lo]


def rmse_vs_track(traj, track_t, track_p):
    """calculate_accuracy_metrics (KF_SensorFusion.ipynb:766-802): RMSE of the Euclidean position
    error of traj [(t, x, y, z, ...)] against the linearly interpolated track."""
    t = np.array([s[0] for s in traj], dtype=float)
    p = np.array([s[1:4] for s in traj], dtype=float)
    ref = np.stack([lerp_extrapolate(t, track_t, track_p[:, k]) for k in range(3)], axis=1)
    e = np.linalg.norm(p - ref, axis=1)
    return np.sqrt(np.mean(e ** 2))


def accuracy_metrics(candidate, ground_truth):
    """kf_workers.calculate_accuracy_metrics (kf_workers.py:1162-1216): the ground-truth section
    inside the candidate's time span (whole track if it has < 2 samples, :1185-1187), interpolated
    at the candidate times; returns (rmse, interpolated positions)."""
    c0, c1 = candidate[0][0], candidate[-1][0]
    sec = [s for s in ground_truth if c0 <= s[0] <= c1]
    if len(sec) < 2:
        sec = ground_truth
    gt_t = np.array([s[0] for s in sec], dtype=float)
    gt_p = np.array([s[1:4] for s in sec], dtype=float)
    ct = np.array([s[0] for s in candidate], dtype=float)
    cp = np.array([s[1:4] for s in candidate], dtype=float)
    interp = np.stack([lerp_extrapolate(ct, gt_t, gt_p[:, k]) for k in range(3)], axis=1)
    e = np.linalg.norm(cp - interp, axis=1)
    return np.sqrt(np.mean(e ** 2)), interp


def product_combo_run(model, combo, x0, P0):
    """One sequence of the product search (KF_SensorFusion.ipynb:896-930): first event dt = 0."""
    x, P = np.array(x0, float).copy(), np.array(P0, float).copy()
    traj, prev = [], None
    for (_, stype, t, sdata) in combo:
        dt = t - prev if prev is not None else 0
        x, P = step(model, x, P, _type_code(stype), dt, event_payload(model, stype, sdata))
        traj.append((t, *x[:6]))
        prev = t
    return traj, x, P


def product_search(model, queues, track_t, track_p, x0=None, P0=None):
    """itertools.product(*queues) in order, strict '<' keeps the first minimum (:932-938)."""
    x0 = np.zeros(model.n) if x0 is None else x0
    P0 = nb_P0() if P0 is None else P0
    best = (float("inf"), None, None, None, None, -1)
    metrics = []
    for ci, combo in enumerate(itertools.product(*queues)):
        traj, x, P = product_combo_run(model, combo, x0, P0)
        mtr = rmse_vs_track(traj, track_t, track_p)
        metrics.append(mtr)
        if mtr < best[0]:
            best = (mtr, traj, combo, x, P, ci)
    return {"selected_sensors": best[2], "final_state": best[3], "final_covariance": best[4],
            "trajectory": best[1], "accuracy_metric": best[0], "combo_index": best[5],
            "metrics": np.array(metrics)}


# ---- sensor_scheduling_algorithm.ipynb:213-291 (a17): exhaustive multi-day schedule tree -------------
def bf_schedule(P, H, z, days, k):
    """Restatement of the notebook's BF_schedule sketch (not runnable as written: it reads undefined
    globals / attributes), loop for loop: all C(G, k) subsets per day, covariance-only update
    `cov_matrix` (:301-376) with R_t = diag(0.01 z_t^2) (:233-234), one global node index per level
    (:246-247), minimum trace over the LAST level (:249-253), parents by divmod (:265-272), lists
    reversed (:275-283).  Sensors are numbered 1..G as cov_matrix expects.  parity: restatement only."""
    G = H.shape[0]
    S_all = list(itertools.combinations(range(1, G + 1), k))
    sets = [[np.array(P, dtype=np.float64)]]
    best = None
    for t in range(1, days + 1):
        R = np.diag(0.01 * np.asarray(z[t - 1], dtype=np.float64) ** 2)
        cur, obj_min = [], np.inf
        for j in range(len(S_all) ** (t - 1)):
            for S_cur in S_all:
                Sig = cov_matrix(list(S_cur), sets[t - 1][j], R, H)
                cur.append(Sig)
                obj = np.trace(Sig)
                if obj < obj_min:
                    obj_min, best = obj, (S_cur, Sig.copy(), len(cur) - 1)
        sets.append(cur)
    S_g, Sig_g = [best[0]], [np.trace(best[1])]
    a, b = divmod(best[2], len(S_all))
    for t in range(days - 1):
        Sig_g.append(np.trace(sets[days - t - 1][a]))
        idx = a
        a, b = divmod(idx, len(S_all))
        S_g.append(S_all[b])
    return Sig_g[::-1], S_g[::-1]
