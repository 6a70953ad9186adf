"""Drop-in mirror of the class defined in the reference's `KF_SensorFusion.ipynb` (cell 2) for the
one hot-path function that exists ONLY there: `run_brute_force_kalman_filter`, the product-of-
queues brute force scored by RMSE against the GPS track (KF_SensorFusion.ipynb:804-954), with its
`calculate_accuracy_metrics` (:766-802) and `get_utm_data`.

The host keeps what the notebook does outside the arithmetic: bucketing the events into one queue
per sampling period (:826-847), the prints, and re-shaping the winner into the notebook's dict.
Every filter step, the track interpolation, the RMSE and the argmin run in libb200kf.so
(b200kf_track_lerp, b200kf_score_products, b200kf_run_batch); there is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np

from . import kf_workers as _kw
from . import streams


class KF_SensorFusion(_kw.KF_SensorFusion):
    def __init__(self, gps_csv_file, imu_csv_file, total_num_sensors=15, max_active_num_sensors=1):
        super().__init__(gps_csv_file, imu_csv_file)
        self.total_num_sensors = total_num_sensors
        self.max_active_num_sensors = max_active_num_sensors

    def get_utm_data(self):
        return self.utm_data

    def _track(self):
        utm = self.get_utm_data()
        t = np.array([e["time"] for e in utm], dtype=np.float64)
        p = np.array([[e["easting"], e["northing"], e["altitude"]] for e in utm], dtype=np.float64)
        return t, p

    # ---- KF_SensorFusion.ipynb:766-802: scalar RMSE against the interpolated GPS track ----------
    def calculate_accuracy_metrics(self, sf_KF_state_with_time):
        utm = self.get_utm_data()
        if not sf_KF_state_with_time or not utm:
            print("State list or UTM data is empty. Cannot calculate accuracy.")
            return None
        eng = self._engine()
        tt, tp = self._track()
        kt = np.array([s[0] for s in sf_KF_state_with_time], dtype=np.float64)
        kp = np.array([s[1:4] for s in sf_KF_state_with_time], dtype=np.float64)
        ref = eng.track_lerp(tt, tp, kt)
        rmse = eng.rmse_vs_track(kp[:, :, None], ref)  # [T, 3, N=1] time-major
        return np.float64(rmse[0].item())

    # ---- KF_SensorFusion.ipynb:826-847 ------------------------------------------------------------
    def _bucket_queues(self, start_idx, end_idx, processing_frequency):
        queues, queue = [], []
        started, previous_time = False, None
        for ev in self.indexed_sensor_data[start_idx:end_idx]:
            if not started and ev[1] == "GPS":
                started, previous_time = True, ev[2]
            if not started:
                continue
            if ev[2] - previous_time < 1 / processing_frequency:
                queue.append(ev)
                continue
            if len(queue) == 0:
                queue.append(ev)
            queues.append(queue)
            queue = []
            previous_time = ev[2]
        return queues

    # ---- KF_SensorFusion.ipynb:804-954 ------------------------------------------------------------
    def run_brute_force_kalman_filter(self, start_idx=0, end_idx=None, overwrite_sampling_freq=None,
                                      max_combos_in_memory=1000):
        """All sequences (one event per sampling period) on the GPU; `max_combos_in_memory` is
        accepted for signature compatibility -- nothing is enumerated in host memory."""
        xt = np.zeros(15)
        Pt = np.diag([1000.0] * 3 + [100.0] * 3 + [100.0] * 3 + [100.0] * 3 + [1000.0] * 3)
        if end_idx is None:
            end_idx = len(self.indexed_sensor_data)
        f = overwrite_sampling_freq if overwrite_sampling_freq is not None else self.processing_frequency
        queues = self._bucket_queues(start_idx, end_idx, f)
        qlen = [len(q) for q in queues]
        log_total = sum(math.log10(l) for l in qlen)
        if log_total > 12:
            print(f"Total combinations: ~1e{int(log_total)} (too large to enumerate in memory)")
        else:
            print(f"Total combinations: {math.prod(qlen)}")
        none = {"selected_sensors": None, "final_state": None, "final_covariance": None, "trajectory": None,
                "accuracy_metric": float("inf")}
        if not queues or not self.get_utm_data():
            return none  # no sequence is ever scored (metric None never beats inf, :932-933)
        if len(queues) > 64 or log_total > 19:
            raise NotImplementedError(f"{len(queues)} windows / ~1e{int(log_total)} sequences: the product-search "
                                      "kernel enumerates at most 64 windows and 2^64 sequences")
        flat = [ev for q in queues for ev in q]
        ps = streams.PackedStream.from_events(flat)
        eng = self._engine()
        tt, tp = self._track()
        ref = eng.track_lerp(tt, tp, ps.times)
        best, _ = eng.score_products(qlen, ps.types, ps.times, ps.payload(self.MODEL_NAME), ref, xt, Pt)
        print(f"Processed {best['n_scored']}/{math.prod(qlen)} combinations...")
        if best["index"] is None:
            return none
        # decode the winner (itertools.product order: last queue fastest, :873-875) and replay it once
        digits, rem = [], best["index"]
        for l in reversed(qlen):
            digits.append(rem % l)
            rem //= l
        digits.reverse()
        off = np.concatenate(([0], np.cumsum(qlen)[:-1]))
        sel = off + np.array(digits)
        combo = tuple(flat[i] for i in sel)
        times = ps.times[sel]
        dts = np.concatenate(([0.0], np.diff(times)))  # first event dt = 0 (:898)
        r = eng.run_batch(ps.types[sel], dts, ps.payload(self.MODEL_NAME, sel)[:, :, None], xt[None], Pt[None],
                          want_agg=False, traj_ncomp=6)
        traj = r["traj"].cpu().numpy()[:, :, 0]
        return {"selected_sensors": combo, "final_state": r["x"].cpu().numpy()[0],
                "final_covariance": r["P"].cpu().numpy()[0],
                "trajectory": [(float(t), *row) for t, row in zip(times, traj)],
                "accuracy_metric": best["score"],
                # extras (not in the notebook): the winner's index and the runner-up RMSE
                "combo_index": best["index"], "runner_up_metric": best["runner_up_score"]}


__all__ = ["KF_SensorFusion"]
