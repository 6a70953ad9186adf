"""Where does a 2^20 search call spend its 45 us?  Times (a) Engine.score_schedules from Python, (b) the bare C-ABI call
through ctypes with prepared arguments, both back to back on one stream, and (c) the same with a synchronize per call.
python tools/search_call_probe.py [--L 20] [--reps 2000]"""
import argparse
import ctypes
import json
import os
import sys
import time
from ctypes import POINTER, c_double, c_uint8, c_void_p

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import search_window, warm_cov  # noqa: E402
from sensorfusion_kalmanfilter_b200 import _lib  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=20)
ap.add_argument("--reps", type=int, default=2000)
a = ap.parse_args()
eng = Engine("CA15")
st, tm = search_window(a.L)
P0 = warm_cov(eng)
t_prev, t_end = 999.995, float(tm[-1])
_, per = eng.score_schedules(st, tm, t_prev, t_end, P0, first=1, count=(1 << 16) - 1, want_per_schedule=True)
thr = float(per[:, 0].median().item())
out = {"L": a.L, "reps": a.reps}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    return {"us_per_call": (time.perf_counter() - t0) / reps * 1e6, "us_host_issue": t_issue / reps * 1e6}


out["python_engine"] = timed(lambda: eng.score_schedules(st, tm, t_prev, t_end, P0, threshold=thr, device_result=True), a.reps)
best_d = torch.zeros((ctypes.sizeof(_lib.Best),), dtype=torch.uint8, device=eng.device)
tp, tmp, Pp = st.ctypes.data_as(POINTER(c_uint8)), tm.ctypes.data_as(POINTER(c_double)), P0.ctypes.data_as(POINTER(c_double))
stream = c_void_p(torch.cuda.current_stream().cuda_stream)
bp = c_void_p(best_d.data_ptr())
fn = eng.lib.b200kf_score_schedules
h = eng._h


def raw():
    fn(h, a.L, tp, tmp, t_prev, t_end, Pp, 0, 1 << a.L, thr, float("nan"), c_void_p(0), bp, 0, stream)


out["bare_c_abi"] = timed(raw, a.reps)


def raw_sync():
    raw()
    torch.cuda.synchronize()


out["bare_c_abi_sync_each"] = timed(raw_sync, min(a.reps, 500))
print(json.dumps(out))
