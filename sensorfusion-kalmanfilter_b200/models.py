"""Model descriptors: the VALUES behind the reference's eight model getters.

The CUDA kernels compile in the state layout and the sparsity of F(dt), H and the diagonal shape
of Q(dt), R for five models (CV6, CA8, BIAS9, CA15, CVBU6); the rate constants are data.
`spec_from_getters` turns a `KF_SensorFusion`-like object into a descriptor by CALLING its getters
(kf_workers.py:493-621 / hw5_2.py:219-311) -- so a subclass that overrides Q or R changes the GPU
results, and one that leaves the compiled family (dense Q, different F sparsity, non-selector H)
raises instead of being silently ignored.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib

MODEL_IDS = {"CV6": _lib.MODEL_CV6, "CA8": _lib.MODEL_CA8, "BIAS9": _lib.MODEL_BIAS9,
             "CA15": _lib.MODEL_CA15, "CVBU6": _lib.MODEL_CVBU6}
DIMS = {"CV6": (6, 3, 3), "CA8": (8, 4, 2), "BIAS9": (9, 3, 3), "CA15": (15, 9, 3), "CVBU6": (6, 3, 3)}  # n, d, m_gps

# chains: groups of state indices that F, H, Q, R never couple (csrc/kf_core.cuh)
CHAINS = {
    "CV6": [(0, 3), (1, 4), (2, 5)],
    "CA8": [(0, 3, 6), (1, 4, 7), (2, 5)],
    "BIAS9": [(0, 3, 6), (1, 4, 7), (2, 5, 8)],
    "CA15": [(0, 6, 12), (1, 7, 13), (2, 8, 14), (3, 9), (4, 10), (5, 11)],
    "CVBU6": [(0, 3), (1, 4), (2, 5)],
}


@dataclass
class ModelSpec:
    name: str
    precision: str = "f64"           # "f64" | "f32"
    q: np.ndarray = field(default=None)
    r_gps: np.ndarray = field(default=None)
    r_imu: np.ndarray = field(default=None)
    accel_coef: float = 0.5

    @property
    def model_id(self):
        return MODEL_IDS[self.name]

    @property
    def n(self):
        return DIMS[self.name][0]

    @property
    def d(self):
        return DIMS[self.name][1]

    @property
    def m_gps(self):
        return DIMS[self.name][2]

    @property
    def imu_is_control(self):
        return self.name in ("BIAS9", "CVBU6")

    def default_P0(self):
        """Cold-start covariance of the reference (kf_workers.py:651, hw5_2.py:317-326)."""
        return np.diag(_P0_DIAG[self.name]).astype(np.float64)

    def to_desc(self) -> _lib.ModelDesc:
        d = _lib.ModelDesc()
        d.model_id = self.model_id
        d.precision = _lib.F32 if self.precision == "f32" else _lib.F64
        for i, v in enumerate(self.q):
            d.q[i] = float(v)
        for i, v in enumerate(self.r_gps):
            d.r_gps[i] = float(v)
        for i, v in enumerate(self.r_imu):
            d.r_imu[i] = float(v)
        d.accel_coef = float(self.accel_coef)
        return d


_P0_DIAG = {
    "CV6": [1e4] * 3 + [1e3] * 3,
    "CA8": [1000, 1000, 100, 100, 100, 100, 1000, 1000],
    "BIAS9": [1e4] * 3 + [1e3] * 3 + [1e4] * 3,
    "CA15": [1e4] * 3 + [1e3] * 3 + [1e3] * 3 + [1e3] * 3 + [1e4] * 3,
    "CVBU6": [1e4] * 3 + [1e3] * 3,
}
_Q = {
    "CV6": [5] * 3 + [1] * 3,
    "CA8": [5, 5, 0.05, 1, 1, 0.1, 2, 2],
    "BIAS9": [5] * 3 + [1] * 3 + [2] * 3,
    "CA15": [5] * 3 + [0.05] * 3 + [1] * 3 + [0.1] * 3 + [2] * 3,
    "CVBU6": [5] * 3 + [1] * 3,
}
_R_IMU = {
    "CV6": [50] * 3 + [10] * 3,
    "CA8": [50, 50, 0.05, 10, 10, 0.1, 100, 100],
    "BIAS9": [0] * 9,
    "CA15": [50] * 3 + [0.05] * 3 + [10] * 3 + [0.1] * 3 + [100] * 3,
    "CVBU6": [0] * 6,
}


def default_spec(name: str, precision: str = "f64") -> ModelSpec:
    """The reference's own constants (kf_workers.py:519-614, hw5_2.py:233-304)."""
    n, _, m = DIMS[name]
    return ModelSpec(name=name, precision=precision, q=np.array(_Q[name], dtype=np.float64),
                     r_gps=np.full(m, 3.0), r_imu=np.array(_R_IMU[name], dtype=np.float64))


def _links(name):
    """(i, j, kind) off-diagonal entries of F: kind 1 -> dt, kind 2 -> accel_coef*dt^2 (signed)."""
    out = []
    for ch in CHAINS[name]:
        sgn1 = {0: 1.0, 1: 1.0}
        sgn2 = 1.0
        if name == "BIAS9":
            sgn1 = {0: 1.0, 1: -1.0}
            sgn2 = -1.0
        if len(ch) >= 2:
            out.append((ch[0], ch[1], 1, sgn1[0]))
        if len(ch) == 3:
            out.append((ch[1], ch[2], 1, sgn1[1]))
            out.append((ch[0], ch[2], 2, sgn2))
    return out


def spec_from_getters(obj, name: str, precision: str = "f64") -> ModelSpec:
    """Probe the getters of a KF_SensorFusion-like object and check they stay inside the compiled
    family of `name`.  Raises NotImplementedError with the offending getter otherwise."""
    n, _, m = DIMS[name]
    dts = (0.005, 0.1, 0.37)
    coef = None
    for dt in dts:
        F = np.asarray(obj.get_state_transition_matrix(dt), dtype=np.float64)
        if F.shape != (n, n):
            raise NotImplementedError(f"get_state_transition_matrix returns {F.shape}, model {name} needs {(n, n)}")
        E = np.eye(n)
        for (i, j, kind, sg) in _links(name):
            if kind == 1:
                E[i, j] = sg * dt
            else:
                c = F[i, j] / (sg * dt * dt)
                coef = c if coef is None else coef
                E[i, j] = sg * coef * dt * dt
        if not np.array_equal(F, E) and not np.allclose(F, E, rtol=1e-15, atol=0.0):
            raise NotImplementedError(
                f"get_state_transition_matrix({dt}) is outside the compiled {name} sparsity "
                "(I + dt*N1 + c*dt^2*N2 on the position/velocity/acceleration chains)")
    if coef is None:
        coef = 0.5
    q = None
    for dt in dts:
        Q = np.asarray(obj.get_process_noise_covariance_matrix(dt), dtype=np.float64)
        if Q.shape != (n, n) or np.count_nonzero(Q - np.diag(np.diag(Q))):
            raise NotImplementedError("get_process_noise_covariance_matrix must be diagonal [n, n]")
        qd = np.diag(Q) / dt
        if q is None:
            q = qd
        elif not np.allclose(qd, q, rtol=1e-12, atol=0.0):
            raise NotImplementedError("get_process_noise_covariance_matrix must be dt * diag(q)")
    Hg = np.asarray(obj.get_gps_observation_matrix(), dtype=np.float64)
    Eg = np.zeros((m, n))
    Eg[np.arange(m), np.arange(m)] = 1.0
    if Hg.shape != (m, n) or not np.array_equal(Hg, Eg):
        raise NotImplementedError(f"get_gps_observation_matrix must select states 0..{m - 1}")
    Rg = np.asarray(obj.get_gps_measurement_noise_covariance_matrix(), dtype=np.float64)
    if Rg.shape != (m, m) or np.count_nonzero(Rg - np.diag(np.diag(Rg))) or np.any(np.diag(Rg) <= 0):
        raise NotImplementedError("get_gps_measurement_noise_covariance_matrix must be a positive diagonal")
    Hi = np.asarray(obj.get_imu_observation_matrix(), dtype=np.float64)
    if Hi.shape != (n, n) or not np.array_equal(Hi, np.eye(n)):
        raise NotImplementedError("get_imu_observation_matrix must be the identity")
    Ri = np.asarray(obj.get_imu_measurement_noise_covariance_matrix(), dtype=np.float64)
    if Ri.shape != (n, n) or np.count_nonzero(Ri - np.diag(np.diag(Ri))) or np.any(np.diag(Ri) <= 0):
        raise NotImplementedError("get_imu_measurement_noise_covariance_matrix must be a positive diagonal")
    # predict_covariance / calculate_kalman_gain: probe that they still are F P F^T + Q and
    # P H^T (H P H^T + R)^-1 on one random SPD matrix.
    rng = np.random.default_rng(0)
    A = rng.normal(size=(n, n))
    Pp = A @ A.T + np.eye(n)
    Fp = np.asarray(obj.get_state_transition_matrix(0.1), dtype=np.float64)
    Qp = np.asarray(obj.get_process_noise_covariance_matrix(0.1), dtype=np.float64)
    if not np.allclose(obj.predict_covariance(Pp, Fp, Qp), Fp @ Pp @ Fp.T + Qp, rtol=1e-10, atol=1e-10):
        raise NotImplementedError("predict_covariance is not F P F^T + Q")
    Kp = obj.calculate_kalman_gain(Pp, Hg, Rg)
    if not np.allclose(Kp, Pp @ Hg.T @ np.linalg.inv(Hg @ Pp @ Hg.T + Rg), rtol=1e-9, atol=1e-9):
        raise NotImplementedError("calculate_kalman_gain is not P H^T (H P H^T + R)^-1")
    return ModelSpec(name=name, precision=precision, q=np.array(q), r_gps=np.diag(Rg).copy(),
                     r_imu=np.diag(Ri).copy(), accel_coef=float(coef))


def chain_compatible(name: str, P: np.ndarray) -> bool:
    """True when P [..., n, n] has exact zeros outside the chains (decoupled kernel valid)."""
    n = DIMS[name][0]
    mask = np.ones((n, n), dtype=bool)
    for ch in CHAINS[name]:
        for i in ch:
            for j in ch:
                mask[i, j] = False
    return not np.any(np.asarray(P)[..., mask] != 0)


def axis_symmetric(name: str, P: np.ndarray) -> bool:
    """True when P [..., n, n] is chain-compatible AND the chains that are the same model on different
    axes (the chains of equal length: x/y/z position chains, roll/pitch/yaw chains) carry bit-identical
    blocks -- the symmetric kernels' validity condition for the covariance (kf_core.cuh SymBlock)."""
    if not chain_compatible(name, P):
        return False
    P = np.asarray(P)
    groups = {}
    for ch in CHAINS[name]:
        groups.setdefault(len(ch), []).append(ch)
    for chains in groups.values():
        ref = P[..., np.ix_(chains[0], chains[0])[0], np.ix_(chains[0], chains[0])[1]]
        for ch in chains[1:]:
            blk = P[..., np.ix_(ch, ch)[0], np.ix_(ch, ch)[1]]
            if np.any(blk != ref):
                return False
    return True


def rates_axis_symmetric(spec: ModelSpec) -> bool:
    """True when the chains that are the same model on different axes also share their rate constants
    (q, r_imu, r_gps) -- the second half of the symmetric kernels' validity condition
    (kf_core.cuh SymBlock::constants_equal)."""
    groups = {}
    for ch in CHAINS[spec.name]:
        groups.setdefault(len(ch), []).append(ch)
    for chains in groups.values():
        c0 = chains[0]
        for ch in chains[1:]:
            for a, b in zip(c0, ch):
                if spec.q[a] != spec.q[b] or spec.r_imu[a] != spec.r_imu[b]:
                    return False
                if a < spec.m_gps and b < spec.m_gps and spec.r_gps[a] != spec.r_gps[b]:
                    return False
    return True
