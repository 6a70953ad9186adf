"""Host stream packer: the reference's `indexed_sensor_data` list -> SoA arrays for the C ABI.

The reference walks a Python list of `(i, 'GPS'|'IMU', t, data)` tuples per call
(kf_workers.py:375-385, :681).  Here the list is packed ONCE into three arrays and every run_*
call slices them; `dt` is computed on the host in float64 exactly as the reference does
(`time - prev_time`, kf_workers.py:682), including its out-of-order rules.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

GPS, IMU, PREDICT_ONLY = 0, 1, 2


class PackedStream:
    """types u8 [E], times f64 [E], data f64 [E, 9]: GPS rows (easting, northing, altitude, 0...),
    IMU rows (roll, pitch, yaw, wx, wy, wz, ax, ay, az) = sensor_data[1:10] (kf_workers.py:367)."""

    def __init__(self, types, times, data):
        self.types = types
        self.times = times
        self.data = data

    def __len__(self):
        return len(self.types)

    @classmethod
    def from_events(cls, events):
        E = len(events)
        types = np.empty(E, dtype=np.uint8)
        times = np.empty(E, dtype=np.float64)
        data = np.zeros((E, 9), dtype=np.float64)
        for k, (_, stype, t, sdata) in enumerate(events):
            times[k] = t
            if stype == "GPS":
                types[k] = GPS
                data[k, 0] = sdata["easting"]
                data[k, 1] = sdata["northing"]
                data[k, 2] = sdata.get("altitude", 0.0)  # the 2-D loader (hw5_2.py:52) keeps no altitude
            else:
                types[k] = IMU
                data[k] = sdata[1:10]
        return cls(types, times, data)

    def payload(self, model_name, sl=slice(None)):
        """[E', d] payload rows of `model_name` (layout documented in include/b200kf.h)."""
        ty = self.types[sl]
        da = self.data[sl]
        if model_name == "CA15":
            return np.ascontiguousarray(da)
        gps = (ty == GPS)[:, None]
        if model_name == "CA8":
            imu = da[:, [2, 5, 6, 7]]  # yaw, wz, ax, ay  (hw5_2.py:362-369)
            g = np.zeros_like(imu)
            g[:, :2] = da[:, :2]
            return np.ascontiguousarray(np.where(gps, g, imu))
        imu = da[:, 6:9]  # CV6 / BIAS9: acceleration only
        return np.ascontiguousarray(np.where(gps, da[:, :3], imu))


_cache: "OrderedDict[int, tuple]" = OrderedDict()
_CACHE_MAX = 4


def pack_events(events) -> PackedStream:
    """Pack an indexed_sensor_data list, cached per list object (id + length + end times) so the
    experiment loops that call run_* hundreds of times on one stream pack it once."""
    key = id(events)
    sig = (len(events), events[0][2] if events else None, events[-1][2] if events else None)
    hit = _cache.get(key)
    if hit is not None and hit[0] == sig:
        _cache.move_to_end(key)
        return hit[1]
    ps = PackedStream.from_events(events)
    _cache[key] = (sig, ps)
    while len(_cache) > _CACHE_MAX:
        _cache.popitem(last=False)
    return ps


def schedule_dts(times, prev_time, skip_advances_prev=True):
    """dt per event under the reference's out-of-order rule.  Returns (keep_idx, dts, last_prev).

    run_kalman_filter_full: `dt = time - prev_time; if dt < 0: prev_time = time; continue`
    (kf_workers.py:682-685) -> skip_advances_prev=True.  The adaptive / no-update loops and the
    brute-force worker `continue` WITHOUT moving previous_time (:1013-1015 assigns a dead variable,
    :38-40) -> skip_advances_prev=False."""
    times = np.asarray(times, dtype=np.float64)
    E = len(times)
    if E == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0), prev_time
    dts = np.empty(E)
    dts[0] = times[0] - prev_time
    if E > 1:
        np.subtract(times[1:], times[:-1], out=dts[1:])
    if not np.any(dts < 0):
        return np.arange(E), dts, float(times[-1])
    keep, out = [], []
    prev = prev_time
    for k in range(E):
        dt = times[k] - prev
        if dt < 0:
            if skip_advances_prev:
                prev = times[k]
            continue
        keep.append(k)
        out.append(dt)
        prev = times[k]
    return np.array(keep, dtype=np.int64), np.array(out, dtype=np.float64), float(prev)
