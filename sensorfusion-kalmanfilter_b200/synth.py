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


def write_config1_csvs(directory, n_imu=4000, n_gps_nan=12, seed=0, gps_csv=None):
    """BASELINE configs[0] inputs: `imu_data.csv` in the schema hw5_1.py:29-31
    writes -- header + 11 columns (time, orientation x/y/z/w, angular_velocity x/y/z,
    linear_acceleration x/y/z), 200 Hz, t_k = 1697739278.7381794 + 0.005 k, quaternion of small Euler
    angles around (-0.0004, -0.0547, -0.0008) rad, w = bias + N(0, 0.002^2), a = bias + N(0, 0.05^2)
    with the biases the reference measured (KF_SensorFusion.ipynb:1325-1328) -- and, unless `gps_csv`
    names the reference's own gps_data.csv, a synthetic `gps_data.csv` in its schema (hw5_1.py:28:
    time, latitude, longitude, altitude; ~10 Hz; leading rows without a fix are 'nan') on a slow
    track around the data set's location.  Values are written with csv.writer so float(entry[k])
    round-trips.  Returns (gps_path, imu_path)."""
    import csv
    import os
    rng = np.random.default_rng(seed)
    t0 = 1697739278.7381794
    t_imu = t0 + 0.005 * np.arange(n_imu)
    rpy = np.array([-0.0004, -0.0547, -0.0008]) + rng.normal(0.0, 0.002, (n_imu, 3))
    cr, sr = np.cos(rpy[:, 0] / 2), np.sin(rpy[:, 0] / 2)
    cp, sp = np.cos(rpy[:, 1] / 2), np.sin(rpy[:, 1] / 2)
    cy, sy = np.cos(rpy[:, 2] / 2), np.sin(rpy[:, 2] / 2)
    quat = np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], axis=1)  # x, y, z, w
    omg = np.array([-0.0016706, -0.00751355, -0.03627881]) + rng.normal(0.0, 0.002, (n_imu, 3))
    acc = np.array([-0.523625, 0.00864235, -9.5308909]) + rng.normal(0.0, 0.05, (n_imu, 3))
    imu_path = os.path.join(directory, "imu_data.csv")
    with open(imu_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["time", "orientation_x", "orientation_y", "orientation_z", "orientation_w",
                    "angular_velocity_x", "angular_velocity_y", "angular_velocity_z",
                    "linear_acceleration_x", "linear_acceleration_y", "linear_acceleration_z"])
        for k in range(n_imu):
            w.writerow([repr(float(t_imu[k]))] + [repr(float(v)) for v in quat[k]] + [repr(float(v)) for v in omg[k]]
                       + [repr(float(v)) for v in acc[k]])
    if gps_csv is not None:
        return gps_csv, imu_path
    n_gps = max(2, int(n_imu * 0.005 / 0.1))
    t_gps = t0 + 0.0234 + 0.1 * np.arange(n_gps) + 1e-4 * rng.random(n_gps)
    s = t_gps - t_gps[0]
    lat = 42.3482 + 1e-6 * s + 2e-7 * rng.normal(size=n_gps)       # ~0.1 m/s northwards, ~2 cm noise
    lon = -71.0924 + 2e-6 * np.sin(0.05 * s) + 2e-7 * rng.normal(size=n_gps)
    alt = -19.5 + 0.01 * s + 0.05 * rng.normal(size=n_gps)
    gps_path = os.path.join(directory, "gps_data.csv")
    with open(gps_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["time", "latitude", "longitude", "altitude"])
        for k in range(n_gps):
            if k < n_gps_nan:
                w.writerow([repr(float(t_gps[k])), "nan", "nan", "nan"])
            else:
                w.writerow([repr(float(t_gps[k])), repr(float(lat[k])), repr(float(lon[k])), repr(float(alt[k]))])
    return gps_path, imu_path


def seed_noise(types, seed, sigma_imu=0.05, var_gps=3.0):
    """Measurement noise of Monte-Carlo seed `seed` for a packed base stream (BASELINE configs[1],
    SURVEY.md 8(d) config 2): GPS rows get N(0, R_GPS = 3) on their three columns, IMU rows N(0, sigma_imu^2)
    on all nine channels.  Returns [E, 9] float64 to ADD to the base stream's data; seed 0 .. S-1 are
    independent `default_rng(seed)` streams, so any subset of seeds can be regenerated on its own."""
    types = np.asarray(types)
    rng = np.random.default_rng(seed)
    out = rng.normal(0.0, sigma_imu, (len(types), 9))
    g = types == 0
    ng = int(g.sum())
    if ng:
        rows = np.zeros((ng, 9))
        rows[:, :3] = rng.normal(0.0, np.sqrt(var_gps), (ng, 3))
        out[g] = rows
    return out


def write_gps_csv(path, time, lat, lon, alt):
    """A `gps_data.csv` (hw5_1.py:28 schema: time, latitude, longitude, altitude; rows without a fix are
    'nan') from float64 columns, written the way hw5_1.py writes it (csv.writer, repr of the floats), so a
    log stored as arrays is rebuilt byte for byte."""
    import csv

    def txt(v):
        return "nan" if np.isnan(v) else repr(float(v))
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["time", "latitude", "longitude", "altitude"])
        for k in range(len(time)):
            w.writerow([repr(float(time[k])), txt(lat[k]), txt(lon[k]), txt(alt[k])])
    return path
