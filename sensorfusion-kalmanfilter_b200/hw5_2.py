"""Drop-in mirror of the reference's `hw5_2.py` filter loops (8-state 2-D model) on the B200 engine.

`KF_SensorFusion.run_kalman_filter()` -> `[(x, y, theta), ...]` (hw5_2.py:313-380) and
`run_dead_reckoning_for_IMU()` (hw5_2.py:382-436).  The loops run in b200kf_run_batch with the
compiled CA8 model; the getters below carry the reference's constants (hw5_2.py:219-304) and may be
overridden (values only -- see models.spec_from_getters).  Plotting is out of scope.
"""
from __future__ import annotations

import numpy as np

from . import geo, streams
from . import kf_workers as _kw
from .kf_workers import _engine_for


class KF_SensorFusion:
    MODEL_NAME = "CA8"

    def __init__(self, gps_csv_file, imu_csv_file):
        self.gps_csv_file = gps_csv_file
        self.imu_csv_file = imu_csv_file
        self.gps_data = []
        self.imu_data = []
        self.utm_data = []

    # ---- CSV front end (hw5_2.py:17-112): the same preparation code as kf_workers.py, without altitude
    load_data_from_csv = _kw.KF_SensorFusion.load_data_from_csv
    load_data = _kw.KF_SensorFusion.load_data
    compute_imu_biases = staticmethod(_kw.KF_SensorFusion.compute_imu_biases)
    quaternion_to_euler = _kw.KF_SensorFusion.quaternion_to_euler
    unbias_imu_data = _kw.KF_SensorFusion.unbias_imu_data
    combine_sensor_data = _kw.KF_SensorFusion.combine_sensor_data

    def gps_to_utm(self):
        """hw5_2.py:29-52 (lat/lon -> UTM restated in geo.py, parity unpinned)."""
        initial_easting = initial_northing = None
        for entry in self.gps_data:
            time, lat_str, lon_str, _ = entry
            if "nan" in lat_str.lower() or "nan" in lon_str.lower():
                continue
            easting, northing, zone_number, zone_letter = geo.from_latlon(float(lat_str), float(lon_str))
            if initial_easting is None or initial_northing is None:
                initial_easting, initial_northing = easting, northing
            self.utm_data.append({"time": float(time), "easting": easting - initial_easting,
                                  "northing": northing - initial_northing, "zone_number": zone_number,
                                  "zone_letter": zone_letter})

    # ---- model getters (hw5_2.py:219-311) ------------------------------------------------------
    def get_state_transition_matrix(self, dt):
        F = np.eye(8)
        F[0, 3] = F[1, 4] = F[2, 5] = F[3, 6] = F[4, 7] = dt
        F[0, 6] = F[1, 7] = 0.5 * dt ** 2
        return F

    def get_process_noise_covariance_matrix(self, dt):
        return np.diag([5 * dt, 5 * dt, 0.05 * dt, 1 * dt, 1 * dt, 0.1 * dt, 2 * dt, 2 * dt])

    def predict_covariance(self, Pt, F, Qt):
        return np.dot(np.dot(F, Pt), F.T) + Qt

    def get_gps_observation_matrix(self):
        H = np.zeros((2, 8), dtype=np.int64)
        H[0, 0] = H[1, 1] = 1
        return H

    def get_imu_observation_matrix(self):
        return np.eye(8, dtype=np.int64)

    def get_gps_measurement_noise_covariance_matrix(self):
        return np.diag([3, 3])

    def get_imu_measurement_noise_covariance_matrix(self):
        return np.diag([50, 50, 0.05, 10, 10, 0.1, 100, 100])

    def calculate_kalman_gain(self, P_next, H, R):
        S = np.dot(np.dot(H, P_next), H.T) + R
        return np.dot(np.dot(P_next, H.T), np.linalg.inv(S))

    def _P0(self):
        return np.diag([1000.0, 1000.0, 100.0, 100.0, 100.0, 100.0, 1000.0, 1000.0])

    def _filter(self, idx, dts):
        ps = streams.pack_events(self.indexed_sensor_data)
        eng = _engine_for(self, self.MODEL_NAME)
        pay = ps.payload(self.MODEL_NAME, idx)
        out = eng.run_batch(ps.types[idx], dts, pay[:, :, None], np.zeros((1, 8)), self._P0()[None],
                            want_agg=False, traj_ncomp=3)
        return out["traj"].cpu().numpy()[:, :, 0]

    # ---- hw5_2.py:313-380 ------------------------------------------------------------------------
    def run_kalman_filter(self):
        """Zero initial state, P0 = diag(1000,1000,100,100,100,100,1000,1000); events before the first
        GPS fix are skipped and that fix is processed with dt = 0."""
        states = [(0, 0, 0)]
        ev = self.indexed_sensor_data
        first = next((i for i, e in enumerate(ev) if e[1] == "GPS"), None)
        if first is None:
            return states
        ps = streams.pack_events(ev)
        idx = np.arange(first, len(ev))
        times = ps.times[idx]
        dts = np.diff(times, prepend=times[0])
        tr = self._filter(idx, dts)
        return states + [(r[0], r[1], r[2]) for r in tr]

    # ---- hw5_2.py:382-436 ------------------------------------------------------------------------
    def run_dead_reckoning_for_IMU(self):
        """IMU events only; the first one is processed with dt = 0."""
        ps = streams.pack_events(self.indexed_sensor_data)
        idx = np.nonzero(ps.types == streams.IMU)[0]
        if len(idx) == 0:
            return []
        times = ps.times[idx]
        dts = np.diff(times, prepend=times[0])
        tr = self._filter(idx, dts)
        return [(r[0], r[1], r[2]) for r in tr]
