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
_PROBES = 64


def _fingerprint(events):
    """Length + the (type, time, first payload values) of up to 64 evenly spaced events: cheap evidence that
    the list still holds what was packed (an in-place edit of the list almost surely moves one of them)."""
    E = len(events)
    if E == 0:
        return (0,)
    step = max(1, E // _PROBES)
    probe = []
    for k in list(range(0, E, step)) + [E - 1]:
        _, stype, t, sdata = events[k]
        if stype == "GPS":
            probe.append((0, t, sdata["easting"], sdata["northing"]))
        else:
            probe.append((1, t, sdata[1], sdata[9]))
    return (E, tuple(probe))


def pack_events(events) -> PackedStream:
    """Pack an indexed_sensor_data list, cached per list OBJECT so the experiment loops that call run_*
    hundreds of times on one stream pack it once.  The cache entry keeps a reference to the list it packed:
    a hit requires `entry_list is events` (CPython reuses the id of a freed list, so an id alone would hand a
    regenerated list of the same shape the old payload) and an unchanged fingerprint of sampled events (an
    in-place edit repacks).  Callers that rewrite single events in place beyond what the 64 probes can see
    must call `invalidate(events)`."""
    key = id(events)
    sig = _fingerprint(events)
    hit = _cache.get(key)
    if hit is not None and hit[0] is events and hit[1] == sig:
        _cache.move_to_end(key)
        return hit[2]
    ps = PackedStream.from_events(events)
    _cache[key] = (events, sig, ps)
    _cache.move_to_end(key)
    while len(_cache) > _CACHE_MAX:
        _cache.popitem(last=False)
    return ps


def invalidate(events=None):
    """Drop the packed form of `events` (or of every list) from the cache."""
    if events is None:
        _cache.clear()
    else:
        _cache.pop(id(events), None)


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
