"""Synthetic GPS/IMU streams in the reference's layouts.

The reference's IMU log (imu_data.csv) is not in its repository and there is no network, so
every stream here is synthesised to the reference's schema:

* `indexed_sensor_data` -- the list `combine_sensor_data` builds (kf_workers.py:375-385):
  `(i, 'GPS', t, {'time','easting','northing','zone_number','zone_letter','altitude'})`
  (kf_workers.py:331) and `(i, 'IMU', t, [str(t), roll, pitch, yaw, wx, wy, wz, ax, ay, az])`
  (kf_workers.py:367), sorted by time.
* packed arrays -- what the C ABI consumes: `types u8[T]`, `dt f64[T]`, `payload f64[T,d,N]`.

Rates, epoch and magnitudes follow the dataset the reference used (200 Hz IMU, 10 Hz GPS, first
fix at t = 1697739552.3362827; KF_SensorFusion.ipynb:1324-1331).
"""
from __future__ import annotations

import numpy as np

T0_FIRST_FIX = 1697739552.3362827  # kf_workers.py:1993


def make_indexed_sensor_data(n_events, seed=0, imu_hz=200.0, gps_hz=10.0, t0=T0_FIRST_FIX,
                             lead_imu=3, jitter=0.2):
    """Build `n_events` time-sorted events.  `lead_imu` IMU events precede the first GPS fix (the
    real log starts with a stationary IMU-only stretch).  Timestamps are float64 epoch seconds
    with per-event jitter so dt varies like a real log."""
    rng = np.random.default_rng(seed)
    imu_dt, gps_dt = 1.0 / imu_hz, 1.0 / gps_hz
    n_gps = max(1, int(np.ceil(n_events * imu_dt / (imu_dt + gps_dt))) + 1)
    n_imu = n_events + lead_imu
    t_imu = t0 - lead_imu * imu_dt + imu_dt * (np.arange(n_imu) + 0.37
                                                + jitter * (rng.random(n_imu) - 0.5))
    t_gps = t0 + gps_dt * np.arange(n_gps) + 1e-4 * jitter * rng.random(n_gps)
    t_gps[0] = t0
    # a smooth ground track: circle-ish motion with slow altitude drift
    def truth(t):
        s = t - t0
        return np.stack([30.0 * np.sin(0.05 * s) + 2.0 * s, 20.0 * (1 - np.cos(0.05 * s)),
                         -32.6 + 0.01 * s], axis=-1)
    def truth_acc(t):
        s = t - t0
        return np.stack([-30.0 * 0.05 ** 2 * np.sin(0.05 * s), 20.0 * 0.05 ** 2 * np.cos(0.05 * s),
                         np.zeros_like(s)], axis=-1)
    gps_pos = truth(t_gps) + rng.normal(0.0, np.sqrt(3.0), (n_gps, 3))
    rpy = np.array([-0.0004, -0.0547, -0.0008]) + rng.normal(0, 0.01, (n_imu, 3))
    omg = rng.normal(0, 0.002, (n_imu, 3))
    acc = truth_acc(t_imu) + rng.normal(0, 0.05, (n_imu, 3))
    merged = []
    for k in range(n_gps):
        t = float(t_gps[k])
        merged.append(("GPS", t, {"time": t, "easting": np.float64(gps_pos[k, 0]),
                                  "northing": np.float64(gps_pos[k, 1]), "zone_number": 19,
                                  "zone_letter": "T", "altitude": float(gps_pos[k, 2])}))
    for k in range(n_imu):
        t = float(t_imu[k])
        merged.append(("IMU", t, [repr(t), np.float64(rpy[k, 0]), np.float64(rpy[k, 1]),
                                  np.float64(rpy[k, 2]), float(omg[k, 0]), float(omg[k, 1]),
                                  float(omg[k, 2]), float(acc[k, 0]), float(acc[k, 1]),
                                  float(acc[k, 2])]))
    merged.sort(key=lambda e: e[1])
    merged = merged[:n_events]
    return [(i, *e) for i, e in enumerate(merged)]


def gps_imu_schedule(n_steps, gps_every=20, dt=0.005, first_dt=0.0):
    """Shared schedule of BASELINE config 3: event k is GPS iff k % gps_every
    == 0, else IMU; dt = 0.005 s (200 Hz), first step dt = 0 as the reference's first event."""
    types = np.where(np.arange(n_steps) % gps_every == 0, 0, 1).astype(np.uint8)
    dts = np.full(n_steps, dt, dtype=np.float64)
    if n_steps:
        dts[0] = first_dt
    return types, dts


def make_payload(model_d, types, n_filters, seed=1234, m_gps=3, dtype=np.float64):
    """Host-side payload [T, d, N] (time-major, filter fastest): GPS rows ~ N(track, 3), IMU rows
    small accelerations / angles ~ N(0, 0.1^2).  Used by parity tests (the throughput bench
    generates the same shape on the device)."""
    rng = np.random.default_rng(seed)
    T = len(types)
    out = rng.normal(0.0, 0.1, (T, model_d, n_filters))
    g = np.asarray(types) == 0
    ng = int(g.sum())
    if ng:
        track = np.cumsum(rng.normal(0, 0.05, (ng, 1, 1)), axis=0)
        gps = track + rng.normal(0.0, np.sqrt(3.0), (ng, m_gps, n_filters))
        rows = out[g]
        rows[:, :m_gps, :] = gps
        rows[:, m_gps:, :] = 0.0
        out[g] = rows
    return np.ascontiguousarray(out.astype(dtype))
