"""Drop-in mirror of the reference's `kf_workers.py` call surface on the B200 engine.

Same names, arguments and return structures as IseanB/SensorFusion-KalmanFilter `kf_workers.py`
(`KF_SensorFusion`, `Scheduler`, `evaluate_combo_chunk_worker`), so the reference's drivers
(kf_workers.py:2251-2396) run unchanged after `import sensorfusion_kalmanfilter_b200.kf_workers as
kf_workers`.  Every filter step, log det, schedule score and cov_matrix runs in the CUDA kernels of
libb200kf.so; this module only does what the reference also does on the host before/after the
loop: pick the start event, compute `dt = t - prev` in float64, and re-shape results into the
reference's lists of tuples.  There is no CPU fallback: without the library or a GPU the run_*
methods raise.

Where the reference's loops are reproduced, the line they follow is cited.  Out of scope (SURVEY.md
section 2): plotting/animation, KalmanFilterLogger, the rosbag extractor.  The CSV front end
(load_data -> gps_to_modified_utm -> compute_imu_biases -> unbias_imu_data -> combine_sensor_data,
kf_workers.py:290-385) is host-side as in the reference; its lat/lon -> UTM step restates the
un-vendored third-party `utm` package (geo.py, parity unpinned).
"""
from __future__ import annotations

import csv
import math
from itertools import combinations

import numpy as np

from . import geo, streams
from .engine import Engine, cov_update
from .models import DIMS, spec_from_getters

_NAME_BY_N = {n: name for name, (n, _, _) in DIMS.items()}
_engines = {}


def _engine_for(getters_obj, name):
    """Engine for the model the getters describe (cached per distinct constants)."""
    spec = spec_from_getters(getters_obj, name)
    key = (name, spec.q.tobytes(), spec.r_gps.tobytes(), spec.r_imu.tobytes(), spec.accel_coef)
    eng = _engines.get(key)
    if eng is None:
        eng = Engine(spec)
        _engines[key] = eng
    return eng


def _torch_mod():
    import torch
    return torch


def _slogdet_signed(P):
    s, l = np.linalg.slogdet(P)  # term 0 is a property of the INPUT matrix (kf_workers.py:32, :675)
    return s * l


class _Getters:
    """Adapter: the `class_args` dict of bound methods (kf_workers.py:1242-1251) as an object."""

    def __init__(self, class_args):
        for k, v in class_args.items():
            if callable(v):
                setattr(self, k, v)


# ------------------------------------------------------------------------------------------
# evaluate_combo_chunk_worker (kf_workers.py:22-97)
# ------------------------------------------------------------------------------------------
def evaluate_combo_chunk_worker(chunk, xt, Pt, class_args, prev_time, target_end_time):
    """Per combo: the full filter over its events (gaps merged into one predict, :37), the log-det
    trace, and the final propagation to target_end_time (:74-82).  Returns the reference's list of
    `(0, traj, combo, x_final, None, log_det, len(combo))` (:86).

    All combos of the chunk run as one batch of b200kf_run_masked: they must be time-ordered
    subsets of one window of at most 64 distinct events (what the reference's driver produces,
    :1330); anything else raises NotImplementedError rather than being computed elsewhere."""
    chunk = list(chunk)
    if not chunk:
        return []
    g = _Getters(class_args)
    n = int(np.asarray(g.get_imu_observation_matrix()).shape[0])
    name = _NAME_BY_N.get(n)
    if name is None:
        raise NotImplementedError(f"no compiled model with {n} states")
    eng = _engine_for(g, name)
    # the shared window: distinct events of the chunk, in time order
    seen = {}
    for combo in chunk:
        for ev in combo:
            seen.setdefault(ev[0], ev)
    window = sorted(seen.values(), key=lambda e: (e[2], e[0]))
    if len(window) > 64:
        raise NotImplementedError("combos of one chunk must come from a window of <= 64 events")
    pos = {ev[0]: i for i, ev in enumerate(window)}
    masks = []
    for combo in chunk:
        last, m = -1, 0
        for ev in combo:
            i = pos[ev[0]]
            if i <= last:
                raise NotImplementedError("each combo must list its events in time order")
            last = i
            m |= 1 << i
        masks.append(m)
    ps = streams.PackedStream.from_events(window)
    out = eng.run_masked(ps.types, ps.times, ps.payload(name), prev_time, target_end_time,
                         np.asarray(xt, dtype=np.float64), np.asarray(Pt, dtype=np.float64), masks,
                         traj_ncomp=min(6, n))
    xf = out["x_final"].cpu().numpy()
    traj = out["traj"].cpu().numpy()
    ld = out["logdet"].cpu().numpy()
    nt = out["n_terms"].cpu().numpy()
    ld0 = _slogdet_signed(np.asarray(Pt, dtype=np.float64))
    results = []
    for r, combo in enumerate(chunk):
        k = abs(int(nt[r]))
        tr = [(float(row[0]), *row[1:]) for row in traj[r, :k]]
        tr[0] = (prev_time, *np.asarray(xt, dtype=np.float64)[:6])
        lds = [ld0] + [v for v in ld[r, 1:k]]
        results.append((0, tr, combo, xf[r].copy(), None, lds, len(combo)))
    return results


# ------------------------------------------------------------------------------------------
# Scheduler (kf_workers.py:99-233)
# ------------------------------------------------------------------------------------------
class Scheduler:
    """Sensor scheduler.  `device` is accepted for signature compatibility; the posterior
    covariance is always computed by the batched CUDA kernel (b200kf_cov_update) -- the reference's
    own device='cuda' branch (:148-169) is the one place it already meant to use the GPU."""

    def cov_matrix(_, S, Sigma_prev, R, H, device="cuda"):  # noqa: N805  (reference signature, :112)
        assert type(S) == list, "S should be a list"
        assert len(S) > 0, "S should not be empty"
        assert type(Sigma_prev) == np.ndarray, "Sigma_prev should be a np.ndarray array"
        assert type(R) == np.ndarray, "R should be a np.ndarray array"
        assert type(H) == np.ndarray, "H should be a np.ndarray array"
        rows = np.array(sorted(S), dtype=np.int32) - 1 if len(S) != R.shape[0] else np.arange(R.shape[0], dtype=np.int32)
        post, _ = cov_update(Sigma_prev[None].astype(np.float64), rows, H.astype(np.float64), R.astype(np.float64))
        return post[0].cpu().numpy()

    def gain(self, measurement=None, S_sigma=None, measurement_cov=(tuple), observation_cov=(tuple), device="cuda"):
        assert type(measurement) != list, "only have one measurement"
        assert S_sigma is not None, "S_sigma should not be None"
        if measurement is None or len(measurement) == 0:
            return 0
        kind = "GPS" if measurement[1] == "GPS" else "IMU"
        if measurement[1] not in ("GPS", "IMU"):
            return None
        return np.trace(self.cov_matrix(S=[1], Sigma_prev=S_sigma, R=measurement_cov[kind],
                                        H=observation_cov[kind], device=device))

    def random_schedule(self, num_measurements=None):
        assert num_measurements is not None, "measurements should not be None"
        return np.random.choice(num_measurements)

    def greedy_schedule(self, measurements=None, S_sigma=None, measurement_cov=None, observation_cov=None,
                        device="cuda"):
        """Index of the queue entry with the LARGEST posterior trace, first wins ties (:195-213).
        The two candidate gains (one per sensor type) come from one batched kernel call each."""
        assert measurements is not None
        gains = {}
        for kind in ("GPS", "IMU"):
            if any(m[1] == kind for m in measurements):
                _, tr = cov_update(np.asarray(S_sigma, dtype=np.float64)[None], np.array([0], dtype=np.int32),
                                   np.asarray(observation_cov[kind], dtype=np.float64),
                                   np.asarray(measurement_cov[kind], dtype=np.float64))
                gains[kind] = float(tr[0].item())
        best_gain, best = -np.inf, None
        for i, m in enumerate(measurements):
            gcur = gains[m[1]]
            if gcur > best_gain:
                best_gain, best = gcur, i
        if best is None:
            raise ValueError("None is not in list")  # what measurements.index(None) raises (:213)
        return best

    def randomized_greedy_schedule(self, total_num_sensors, device="cuda"):
        return None

    def BF_schedule(self, z, days, device="cuda", *, P=None, H=None, total_gauges=None, selected_gauges=None):
        """The exhaustive multi-day schedule tree sketched in sensor_scheduling_algorithm.ipynb:213-291
        (a17): every day one of the C(G, k) sensor subsets is used for a covariance-only measurement
        update (`cov_matrix`, no dynamics between days), day t's noise is R_t = diag(0.01 z_t^2) (:233-234),
        all |S_all|^t posteriors of a level are kept, the leaf with the smallest trace wins (:249-253,
        strict '<': first minimum) and its path is back-tracked through the parents (:265-272).

        The sketch is not runnable (it reads undefined globals `total_gauges`, `selected_gauges` and
        attributes `self.P`, `self.H`, `self.S`, `self.A0_day`, `self.Q0`): P, H, G and k are arguments
        here, `z` [days, G] is taken as already de-trended (:232), and sensors are numbered 1..G as
        `cov_matrix` expects (:118; the sketch's range(total_gauges) would wrap sensor 0 to the last row).
        Each level is |S_all| launches of the batched kernel b200kf_cov_update over all nodes of the
        previous level.  Returns (Sigma_g_list, S_g_list, total_time_list) as the sketch does: the trace
        after each day along the best path, the subsets chosen, cumulative seconds per day."""
        import time
        from itertools import combinations as _comb
        torch = _torch_mod()
        P = np.asarray(self.P if P is None else P, dtype=np.float64)
        H = np.asarray(self.H if H is None else H, dtype=np.float64)
        G = H.shape[0] if total_gauges is None else total_gauges
        k = selected_gauges
        S_all = list(_comb(range(1, G + 1), k))
        nS = len(S_all)
        z = np.asarray(z, dtype=np.float64)
        sigma = torch.as_tensor(P[None]).cuda()
        level_traces, total_time_list, total_time = [], [], 0.0
        for t in range(1, days + 1):
            t0 = time.perf_counter()
            R = np.diag(0.01 * z[t - 1] ** 2)
            posts, traces = [], []
            for S_cur in S_all:
                rows = np.array(S_cur, dtype=np.int32) - 1
                post, tr = cov_update(sigma, rows, H, R)
                posts.append(post)
                traces.append(tr)
            # node j of the previous level expanded by subset i gets index j * |S_all| + i (:246-247)
            sigma = torch.stack(posts, dim=1).reshape(-1, P.shape[0], P.shape[0])
            level_traces.append(torch.stack(traces, dim=1).reshape(-1).cpu().numpy())
            total_time += time.perf_counter() - t0
            total_time_list.append(total_time)
        best_idx = int(np.argmin(level_traces[-1]))  # first minimum, as the strict '<' of :249
        S_g_list = [S_all[best_idx % nS]]
        Sigma_g_list = [float(level_traces[-1][best_idx])]
        a = best_idx // nS
        for t in range(days - 1):  # walk up: the parent's trace, then the subset that produced the parent
            Sigma_g_list.append(float(level_traces[days - t - 2][a]))
            S_g_list.append(S_all[a % nS])
            a //= nS
        return Sigma_g_list[::-1], S_g_list[::-1], total_time_list


# ------------------------------------------------------------------------------------------
# KF_SensorFusion (kf_workers.py:277-1392)
# ------------------------------------------------------------------------------------------
class KF_SensorFusion:
    MODEL_NAME = "CA15"

    def __init__(self, gps_csv_file, imu_csv_file):
        self.gps_csv_file = gps_csv_file
        self.imu_csv_file = imu_csv_file
        self.gps_data = []
        self.imu_data = []
        self.utm_data = []
        self.scheduler = Scheduler()
        self.processing_frequency = None

    def set_processing_frequency(self, frequency):
        self.processing_frequency = frequency

    # ---- data preparation (host side, as the reference; kf_workers.py:290-425) -------------
    def load_data_from_csv(self, filename, has_header=True):
        data = []
        with open(filename, newline="") as csvfile:
            reader = csv.reader(csvfile)
            if has_header:
                next(reader)
            for row in reader:
                data.append(row)
        return data

    def load_data(self):
        self.gps_data = self.load_data_from_csv(self.gps_csv_file)
        self.imu_data = self.load_data_from_csv(self.imu_csv_file)

    def gps_to_modified_utm(self):
        """kf_workers.py:304-332: UTM eastings / northings relative to the first valid fix, plus the
        altitude; rows with a NaN field are dropped.  `utm.from_latlon` is restated in geo.py (the
        third-party package is not vendored by the reference: parity for this step is unpinned)."""
        initial_easting = initial_northing = None
        for entry in self.gps_data:
            time, lat_str, lon_str, alt_str = entry
            if "nan" in lat_str.lower() or "nan" in lon_str.lower() or "nan" in alt_str.lower():
                continue
            easting, northing, zone_number, zone_letter = geo.from_latlon(float(lat_str), float(lon_str))
            if initial_easting is None or initial_northing is None:
                initial_easting, initial_northing = easting, northing
            self.utm_data.append({"time": float(time), "easting": easting - initial_easting,
                                  "northing": northing - initial_northing, "zone_number": zone_number,
                                  "zone_letter": zone_letter, "altitude": float(alt_str)})

    @staticmethod
    def compute_imu_biases(gps_data, imu_data):
        first_valid_index = next((i for i, entry in enumerate(gps_data) if "nan" not in entry[1].lower()), None)
        if first_valid_index is None:
            print("Warning: No valid GPS data found. Cannot compute IMU biases.")
            return None, None
        stationary = imu_data[:first_valid_index]
        w = np.array([[float(e[5]), float(e[6]), float(e[7])] for e in stationary])
        a = np.array([[float(e[8]), float(e[9]), float(e[10])] for e in stationary])
        return np.mean(w, axis=0), np.mean(a, axis=0), first_valid_index

    def quaternion_to_euler(self, x, y, z, w):
        roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
        sinp = 2 * (w * y - z * x)
        pitch = np.pi / 2 * np.sign(sinp) if abs(sinp) >= 1 else np.arcsin(sinp)
        yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
        return roll, pitch, yaw

    def unbias_imu_data(self, angular_velocity_bias, linear_acceleration_bias):
        out = []
        for entry in self.imu_data:
            w = np.array([float(entry[5]), float(entry[6]), float(entry[7])]) - angular_velocity_bias
            a = np.array([float(entry[8]), float(entry[9]), float(entry[10])]) - linear_acceleration_bias
            roll, pitch, yaw = self.quaternion_to_euler(float(entry[1]), float(entry[2]), float(entry[3]),
                                                        float(entry[4]))
            out.append(entry[:1] + [roll, pitch, yaw] + w.tolist() + a.tolist() + entry[11:])
        self.unbias_imu_data = out

    def combine_sensor_data(self):
        combined = [("GPS", float(g["time"]), g) for g in self.utm_data]
        combined += [("IMU", float(e[0]), e) for e in self.unbias_imu_data]
        combined.sort(key=lambda x: x[1])
        self.indexed_sensor_data = [(i, *d) for i, d in enumerate(combined)]

    # ---- the eight model getters (kf_workers.py:493-621); overriding them changes the GPU model
    def get_state_transition_matrix(self, dt):
        F = np.eye(15)
        for ax in range(3):
            F[ax, 6 + ax] = dt
            F[ax, 12 + ax] = 0.5 * dt ** 2
            F[3 + ax, 9 + ax] = dt
            F[6 + ax, 12 + ax] = dt
        return F

    def get_process_noise_covariance_matrix(self, dt):
        return np.diag([5 * dt] * 3 + [0.05 * dt] * 3 + [1 * dt] * 3 + [0.1 * dt] * 3 + [2 * dt] * 3)

    def predict_covariance(self, Pt, F, Qt):
        return np.dot(np.dot(F, Pt), F.T) + Qt

    def get_gps_observation_matrix(self):
        H = np.zeros((3, 15), dtype=np.int64)
        H[0, 0] = H[1, 1] = H[2, 2] = 1
        return H

    def get_imu_observation_matrix(self):
        return np.eye(15, dtype=np.int64)

    def get_gps_measurement_noise_covariance_matrix(self):
        return np.diag([3, 3, 3])

    def get_imu_measurement_noise_covariance_matrix(self):
        return np.diag([50] * 3 + [0.05] * 3 + [10] * 3 + [0.1] * 3 + [100] * 3)

    def calculate_kalman_gain(self, P_next, H, R):
        S = np.dot(np.dot(H, P_next), H.T) + R
        return np.dot(np.dot(P_next, H.T), np.linalg.inv(S))

    # ---- helpers -----------------------------------------------------------------------------
    def _engine(self):
        return _engine_for(self, self.MODEL_NAME)

    def _default_P0(self):
        return np.diag([10000] * 3 + [1000] * 3 + [1000] * 3 + [1000] * 3 + [10000] * 3)

    def _clip(self, start_idx, end_idx):
        if start_idx is None or start_idx < 0:
            start_idx = 0
            print("Overriding start_idx to 0.")
        if end_idx is None or end_idx > len(self.indexed_sensor_data):
            end_idx = len(self.indexed_sensor_data)
            print("Overriding end_idx to the length of indexed_sensor_data.")
        return start_idx, end_idx

    def _first_gps(self, start_idx, stop):
        """First-GPS initialisation (kf_workers.py:655-666): (x0, t0, absolute index) or None."""
        xt = np.zeros(15)
        for i, (_, stype, time, sdata) in enumerate(self.indexed_sensor_data[start_idx:stop]):
            if stype == "GPS":
                xt[0], xt[1], xt[2] = sdata["easting"], sdata["northing"], sdata["altitude"]
                return xt, time, start_idx + i
        return None

    def _run(self, idx, dts, xt, Pt, *, types=None, gate=None, want_used=False, want_ptraj=False):
        """Filter the events `idx` (absolute indices) with time steps `dts` on the GPU."""
        ps = streams.pack_events(self.indexed_sensor_data)
        eng = self._engine()
        ty = ps.types[idx] if types is None else types
        pay = ps.payload(self.MODEL_NAME, idx)
        out = eng.run_batch(ty, dts, pay[:, :, None], np.asarray(xt, dtype=np.float64)[None],
                            np.asarray(Pt, dtype=np.float64)[None], want_agg=False, traj_ncomp=6,
                            want_logdet=True, want_used=want_used, want_ptraj=want_ptraj, gate_threshold=gate)
        traj = out["traj"].cpu().numpy()[:, :, 0]
        ld = out["logdet"].cpu().numpy()[:, 0]
        times = ps.times[idx]
        states = [(float(t), *row) for t, row in zip(times, traj)]
        res = {"states": states, "logdet": [v for v in ld], "x": out["x"].cpu().numpy()[0],
               "P": out["P"].cpu().numpy()[0], "times": times}
        if want_used:
            res["used"] = out["used"].cpu().numpy()[:, 0].astype(bool)
        if want_ptraj:
            res["ptraj"] = out["ptraj"].cpu().numpy()[:, :, 0].reshape(-1, 15, 15)
        return res

    # ---- run_kalman_filter_full (kf_workers.py:623-728) ----------------------------------------
    def run_kalman_filter_full(self, start_idx=None, end_idx=None, initial_pt=None, initial_state=None,
                               print_output=False):
        if not hasattr(self, "indexed_sensor_data") or not self.indexed_sensor_data:
            print("Sensor data has not been combined or is empty.")
            return [], [], []
        start_idx, end_idx = self._clip(start_idx, end_idx)
        xt = np.zeros(15)
        if initial_pt is not None and initial_state is not None:
            Pt = np.asarray(initial_pt, dtype=np.float64)
            xt[0:6] = initial_state[1:7]
            prev_time = initial_state[0]
            first = start_idx
        else:
            Pt = self._default_P0().astype(np.float64)
            init = self._first_gps(start_idx, end_idx + 1)
            if init is None:
                print("No GPS data found to initialize the filter.")
                return [], [], []
            xt, prev_time, first = init
        states = [(prev_time, *xt[:6])]
        logdets = [np.linalg.slogdet(Pt)[1]]
        ps = streams.pack_events(self.indexed_sensor_data)
        keep, dts, prev_out = streams.schedule_dts(ps.times[first:end_idx], prev_time, skip_advances_prev=True)
        if len(keep):
            r = self._run(first + keep, dts, xt, Pt)
            states += r["states"]
            logdets += r["logdet"]
            Pt = r["P"]
        self._ground_truth = states
        self._ground_truth_cov = None  # write-only in the reference (:724); not materialised
        if print_output:
            print("---------------------------- Full Kalman Filter ----------------------------\n"
                  f" Processed {len(keep)} measurements out of {end_idx - first} in total, from index {first} to {end_idx}")
        return states, logdets, Pt, prev_out

    def get_GT(self):
        if hasattr(self, "_ground_truth"):
            return self._ground_truth
        print("Ground truth not computed yet. Please run run_kalman_filter_full() first.")
        return None

    # ---- run_kalman_filter (legacy, kf_workers.py:738-824) -------------------------------------
    def run_kalman_filter(self, start_idx, end_idx):
        xt = np.zeros(15)
        Pt = self._default_P0().astype(np.float64)
        states = [(0, xt[0], xt[1], xt[2], xt[3], xt[4], xt[5])]
        covs = [Pt.copy()]
        ev = self.indexed_sensor_data[start_idx:end_idx]
        first = next((i for i, e in enumerate(ev) if e[1] == "GPS"), None)
        if first is None:
            return states, covs
        ps = streams.pack_events(self.indexed_sensor_data)
        lo = (start_idx if start_idx >= 0 else len(self.indexed_sensor_data) + start_idx) + first
        idx = np.arange(lo, lo + len(ev) - first)
        times = ps.times[idx]
        dts = np.diff(times, prepend=times[0])  # first GPS: dt = 0 (:771-775); no dt<0 rule here
        r = self._run(idx, dts, xt, Pt, want_ptraj=True)
        states += r["states"]
        covs += [P for P in r["ptraj"]]
        return states, covs

    # ---- run_kalman_filter_scheduled (kf_workers.py:826-957) -----------------------------------
    def run_kalman_filter_scheduled(self, start_idx=None, end_idx=None, initial_pt=None, initial_state=None,
                                    selection_method=None, print_output=False):
        start_idx, end_idx = self._clip(start_idx, end_idx)
        if selection_method not in ["random", "greedy"]:
            print("Invalid selection_method. Choose either 'random' or 'greedy'.")
            return None
        xt = np.zeros(15)
        if initial_state is not None:
            Pt = np.asarray(initial_pt, dtype=np.float64)
            xt[0:6] = initial_state[1:7]
            first = start_idx
            previous_time = initial_state[0]
        else:
            Pt = self._default_P0().astype(np.float64)
            init = self._first_gps(start_idx, end_idx)
            if init is None:
                print("WARN: No GPS measurement found in the sensor subset.")
                return None, None
            xt, previous_time, first = init
        states = [(previous_time, *xt[:6])]
        logdets = [np.linalg.slogdet(Pt)[1]]
        if end_idx == -1:
            end_idx = len(self.indexed_sensor_data)
        ps = streams.pack_events(self.indexed_sensor_data)
        lo = first + 1
        types, times = ps.types[lo:end_idx], ps.times[lo:end_idx]
        period = 1 / self.processing_frequency
        if selection_method == "greedy":
            sel, _ = self._engine().select_greedy(types, times, [period], [previous_time], Pt[None])
            sel = sel[0].astype(np.int64)
        else:  # random: window bookkeeping only, no filter arithmetic (:877-893)
            sel, prev, q0 = [], previous_time, None
            for e in range(len(times)):
                if times[e] - prev < period:
                    if q0 is None:
                        q0 = e
                    continue
                if q0 is None:
                    q0 = e
                    qlen = 1
                else:
                    qlen = e - q0
                pick = q0 + int(self.scheduler.random_schedule(num_measurements=qlen))
                sel.append(pick)
                prev = times[pick]
                q0 = None
            sel = np.array(sel, dtype=np.int64)
        if len(sel):
            st = times[sel]
            dts = np.diff(st, prepend=previous_time)  # dt = selected_time - previous_time (:914)
            r = self._run(lo + sel, dts, xt, Pt)
            states += r["states"]
            logdets += r["logdet"]
            Pt = r["P"]
        if print_output:
            print(f"Scheduled KF processed {len(sel)} measurements from index {first + 1} to {end_idx}")
        return states, logdets, Pt

    # ---- adaptive threshold / no update (kf_workers.py:959-1159) -------------------------------
    def _adaptive(self, start_idx, end_idx, R_threshold, initial_pt, initial_state, never_update, print_output):
        start_idx, end_idx = self._clip(start_idx, end_idx)
        if R_threshold is None:
            R_threshold = -float("inf")
            print("Overriding R_threshold to - infinity (i.e. max information utilization).")
        measurement_times = []
        xt = np.zeros(15)
        if initial_pt is not None and initial_state is not None:
            Pt = np.asarray(initial_pt, dtype=np.float64)
            xt[0:6] = initial_state[1:7]
            previous_time = initial_state[0]
            first = start_idx
        else:
            Pt = self._default_P0().astype(np.float64)
            init = self._first_gps(start_idx, end_idx)
            if init is None:
                print("WARN: No GPS measurement found in the sensor subset.")
                return None
            xt, previous_time, first = init
            measurement_times.append(previous_time)
        states = [(previous_time, *xt[:6])]
        logdets = [np.linalg.slogdet(Pt)[1]]
        ps = streams.pack_events(self.indexed_sensor_data)
        # dt < 0: `continue` without moving previous_time (the reference assigns a dead variable, :1013-1015)
        keep, dts, prev_out = streams.schedule_dts(ps.times[first:end_idx], previous_time, skip_advances_prev=False)
        if len(keep):
            idx = first + keep
            if never_update:
                r = self._run(idx, dts, xt, Pt, types=np.full(len(idx), streams.PREDICT_ONLY, dtype=np.uint8))
            else:
                gate = R_threshold if R_threshold != -float("inf") else -1.0e300
                r = self._run(idx, dts, xt, Pt, gate=gate, want_used=True)
                measurement_times += [float(t) for t in r["times"][r["used"]]]
            states += r["states"]
            logdets += r["logdet"]
            Pt = r["P"]
            previous_time = prev_out
        if print_output:
            print("---------------------------- Adaptive Kalman Filter ----------------------------\n"
                  f" Processed {len(keep)} measurements out of {end_idx - first} in total, from index {first} to {end_idx}")
        return states, logdets, Pt, previous_time, measurement_times

    def run_adaptive_threshold_kalman_filter(self, start_idx=None, end_idx=None, R_threshold=None, initial_pt=None,
                                             initial_state=None, print_output=False):
        return self._adaptive(start_idx, end_idx, R_threshold, initial_pt, initial_state, False, print_output)

    def run_no_update_kalman_filter(self, start_idx=None, end_idx=None, R_threshold=None, initial_pt=None,
                                    initial_state=None, print_output=False):
        return self._adaptive(start_idx, end_idx, R_threshold, initial_pt, initial_state, True, print_output)

    # ---- accuracy metric (kf_workers.py:1162-1216) ----------------------------------------------
    def calculate_accuracy_metrics(self, candidate_trajectory):
        if not candidate_trajectory:
            print("Candidate trajectory is empty. Cannot calculate accuracy.")
            return None
        if not (hasattr(self, "_ground_truth") and self._ground_truth):
            print("Ground truth trajectory is not available. Run run_kalman_filter_full() first.")
            return None
        c0, c1 = candidate_trajectory[0][0], candidate_trajectory[-1][0]
        gt = self._ground_truth
        sec = [s for s in gt if c0 <= s[0] <= c1]
        if len(sec) < 2:
            sec = gt
        gt_t = np.array([s[0] for s in sec], dtype=np.float64)
        gt_p = np.array([s[1:4] for s in sec], dtype=np.float64)
        ct = np.array([s[0] for s in candidate_trajectory], dtype=np.float64)
        cp = np.array([s[1:4] for s in candidate_trajectory], dtype=np.float64)
        # interpolation and RMSE on the device (b200kf_track_lerp / b200kf_rmse_vs_track); the
        # per-sample error arrays the reference also returns are differences of those outputs
        eng = self._engine()
        ref_d = eng.track_lerp(gt_t, gt_p, ct)
        rmse = np.float64(eng.rmse_vs_track(cp[:, :, None], ref_d)[0].item())
        interp = ref_d.cpu().numpy()
        err = cp - interp
        eu = np.linalg.norm(err, axis=1)
        return {"total_position_rmse": rmse, "position_errors": err, "euclidean_errors": eu,
                "candidate_times": ct, "candidate_positions": cp, "ground_truth_interp": interp,
                "gt_start_time": c0, "gt_end_time": c1}

    # ---- brute-force schedule search (kf_workers.py:1218-1392) ----------------------------------
    def run_brute_force_kalman_filter_no_sampling_min_usage(self, start_idx=0, end_idx=None, R_threshold=None,
                                                            initial_pt=None, initial_state=None,
                                                            max_combos_in_memory=10000):
        if R_threshold is None:
            raise ValueError("R_threshold must be specified for brute force KF.")
        start_idx, end_idx = self._clip(start_idx, end_idx)
        xt = np.zeros(15)
        if initial_pt is not None and initial_state is not None:
            Pt = np.asarray(initial_pt, dtype=np.float64)
            xt[0:6] = initial_state[1:7]
            prev_time = initial_state[0]
            sensor_measurements = list(self.indexed_sensor_data[start_idx:end_idx])
        else:
            Pt = self._default_P0().astype(np.float64)
            init = self._first_gps(start_idx, end_idx + 1)
            sensor_measurements = []
            if init is not None:
                xt, prev_time, first = init
                sensor_measurements = list(self.indexed_sensor_data[first:end_idx + 1])
            if len(sensor_measurements) == 0:
                print("No sensor measurements found after GPS initialization.")
                return None
        last_measurement_time = self.indexed_sensor_data[end_idx - 1][2]
        n = len(sensor_measurements)
        print(f"Total combinations: {2 ** n} (2^{n})")
        print(f"Target end time: {last_measurement_time}")
        if n > 32:
            raise NotImplementedError(f"window of {n} events: the schedule-search kernel enumerates at most 2^32 "
                                      "subsets (the reference itself cannot finish 2^25)")
        ps = streams.PackedStream.from_events(sensor_measurements)
        eng = self._engine()
        best, _ = eng.score_schedules(ps.types, ps.times, prev_time, last_measurement_time, Pt,
                                      threshold=R_threshold, initial_logdet=_slogdet_signed(Pt))
        mask = best["first_feasible_index"]
        n_rescored = 0
        if best["n_near_threshold"] > 0:
            # epsilon guard: the reference compares max(log_det) < R_threshold strictly
            # (:1353); a schedule whose max sits within 1e-9 of the threshold could be decided the other way by a
            # 1e-12 rounding difference, so exactly those are re-scored in the reference's own arithmetic
            mask, n_rescored = self._resolve_near_threshold(eng, ps, sensor_measurements, prev_time,
                                                            last_measurement_time, Pt, R_threshold)
        if mask is None:
            print("\nBrute force search complete. No combination met the R_threshold.")
            return None
        sel = [i for i in range(n) if (mask >> i) & 1]
        combo = tuple(sensor_measurements[i] for i in sel)
        out = eng.run_masked(ps.types, ps.times, ps.payload(self.MODEL_NAME), prev_time, last_measurement_time, xt, Pt,
                             [mask], traj_ncomp=6)
        k = abs(int(out["n_terms"][0].item()))
        traj = out["traj"].cpu().numpy()[0, :k]
        ld = out["logdet"].cpu().numpy()[0, :k]
        tr = [(float(row[0]), *row[1:]) for row in traj]
        tr[0] = (prev_time, *xt[:6])
        print(f"\nFound valid combination with {len(sel)} measurements meeting R_threshold!!!!")
        return {"selected_sensors": combo, "final_state": out["x_final"].cpu().numpy()[0], "final_covariance": None,
                "trajectory": tr, "accuracy_metric": 0,
                "log_determinants": [_slogdet_signed(Pt)] + [v for v in ld[1:]],
                "num_measurements_used": len(sel),
                # extras (not in the reference): how many feasibility decisions sat within 1e-9 of the
                # threshold, and the winner as a bitmask over the window
                "n_near_threshold": best["n_near_threshold"], "n_rescored_on_host": n_rescored, "selected_mask": mask}

    # ---- epsilon guard of the selection (kf_workers.py:1353-1358) -------------------------------
    def _combo_logdets_host(self, combo, Pt, prev_time, target_end_time):
        """log_det trace of ONE combo in the reference's numpy arithmetic, through this instance's own
        (overridable) getters -- the covariance half of the worker, kf_workers.py:32-82.  Used only to
        settle feasibility decisions the GPU reports as undecided to rounding."""
        P = np.array(Pt, dtype=np.float64)
        I = np.eye(P.shape[0])
        log_det = [_slogdet_signed(P)]
        cur = prev_time
        for (_, stype, t, _sdata) in combo:
            dt = t - cur
            if dt < 0:
                continue
            F = self.get_state_transition_matrix(dt)
            P = self.predict_covariance(P, F, self.get_process_noise_covariance_matrix(dt))
            if stype == "GPS":
                H, R = self.get_gps_observation_matrix(), self.get_gps_measurement_noise_covariance_matrix()
            else:
                H, R = self.get_imu_observation_matrix(), self.get_imu_measurement_noise_covariance_matrix()
            K = self.calculate_kalman_gain(P, H, R)
            P = np.dot(I - np.dot(K, H), P)
            cur = t
            log_det.append(_slogdet_signed(P))
        if cur < target_end_time - 1e-8:
            dt = target_end_time - cur
            F = self.get_state_transition_matrix(dt)
            P = self.predict_covariance(P, F, self.get_process_noise_covariance_matrix(dt))
            log_det.append(_slogdet_signed(P))
        return log_det

    def _resolve_near_threshold(self, eng, ps, sensor_measurements, prev_time, target_end_time, Pt, R_threshold,
                                chunk_bits=22):
        """First feasible schedule in the reference's enumeration order when some decisions are within
        1e-9 of the threshold: min key over (schedules the GPU finds feasible with a margin) and (undecided
        schedules that are feasible in the reference's arithmetic).  Returns (mask or None, number re-scored)."""
        torch = _torch_mod()
        n = len(sensor_measurements)
        tol = 1e-9 * max(1.0, abs(R_threshold))
        ld0 = _slogdet_signed(Pt)
        # (1) first schedule that is feasible with a margin: the same search against threshold - tol
        sure, _ = eng.score_schedules(ps.types, ps.times, prev_time, target_end_time, Pt,
                                      threshold=R_threshold - tol, initial_logdet=ld0)
        sure_key = sure["first_feasible_key"] if sure["first_feasible_index"] is not None else None
        # (2) the undecided ones, from the per-schedule maxima
        total = 1 << n
        cand = []
        for first in range(0, total, 1 << chunk_bits):
            count = min(1 << chunk_bits, total - first)
            _, per = eng.score_schedules(ps.types, ps.times, prev_time, target_end_time, Pt, first=first, count=count,
                                         threshold=R_threshold, initial_logdet=ld0, want_per_schedule=True)
            near = torch.nonzero((per[:, 0] - R_threshold).abs() <= tol).flatten()
            cand += [first + int(i) for i in near.cpu().tolist() if first + int(i) != 0]
            del per
        cand.sort(key=lambda m: _lex_key(m, n))
        n_rescored = 0
        for m in cand:  # ascending enumeration order: the first feasible one ahead of `sure` wins
            if sure_key is not None and _lex_key(m, n) > sure_key:
                break
            combo = tuple(sensor_measurements[i] for i in range(n) if (m >> i) & 1)
            n_rescored += 1
            if max(self._combo_logdets_host(combo, Pt, prev_time, target_end_time)) < R_threshold:
                return m, n_rescored
        return sure["first_feasible_index"], n_rescored


def _lex_key(mask, L):
    """Position of subset `mask` in the reference's enumeration (cardinality ascending, then
    itertools.combinations order, kf_workers.py:1325-1336) as a sortable key -- the device's lex_key."""
    rev = int(format(mask, f"0{L}b")[::-1], 2) if L else 0
    return (bin(mask).count("1") << 32) | (((1 << L) - 1) - rev)


def _lerp_extrapolate(x, xp, fp):
    """scipy interp1d(kind='linear', fill_value='extrapolate') for sorted xp (kf_workers.py:1195-1197),
    in scipy's own two-weight form (interp1d._call_linear)."""
    if len(xp) == 1:
        return np.full_like(x, fp[0], dtype=np.float64)
    hi = np.clip(np.searchsorted(xp, x), 1, len(xp) - 1)
    lo = hi - 1
    return ((x - xp[lo]) / (xp[hi] - xp[lo])) * fp[hi] + ((xp[hi] - x) / (xp[hi] - xp[lo])) * fp[lo]


def find_start_idx_for_time_offset(sensor_fusion, target_seconds):
    """Index of the first event at or after `target_seconds` past the data set's first fix
    (kf_workers.py:1985-2003; 1697739552.3362827 is the reference's hard-coded first-fix epoch, :1993).
    None -- with the reference's messages -- when the stream is missing or never reaches the target."""
    if not hasattr(sensor_fusion, "indexed_sensor_data") or not sensor_fusion.indexed_sensor_data:
        print("Sensor data not available. Run combine_sensor_data() first.")
        return None
    first_timestamp = 1697739552.3362827
    target_timestamp = first_timestamp + target_seconds
    times = streams.pack_events(sensor_fusion.indexed_sensor_data).times  # packed once per stream
    hit = np.flatnonzero(times >= target_timestamp)  # the first such event, sorted stream or not (the reference scans)
    if hit.size:
        i = int(hit[0])
        print(f"Found index {i} at time {times[i]:.3f}s (offset: {times[i] - first_timestamp:.3f}s)")
        return i
    print(f"Target time {target_seconds}s not found in data")
    return None


__all__ = ["KF_SensorFusion", "Scheduler", "evaluate_combo_chunk_worker", "find_start_idx_for_time_offset",
           "combinations", "math"]
