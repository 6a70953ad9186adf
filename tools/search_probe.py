"""One brute-force schedule search of a 2^L window on one GPU, for ncu (FP64 instruction counts and the
full-section capture of score_tree_kernel that bench.py's search roofline cites):

    python tools/search_probe.py --L 26 [--reps 3]

Same window, P0 and threshold rule as bench.py's `search` / `search_large` legs.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import search_window, warm_cov  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
eng = Engine("CA15")
st, tm = search_window(a.L)
P0 = warm_cov(eng)
_, per = eng.score_schedules(st, tm, 999.995, float(tm[-1]), P0, first=1, count=(1 << 16) - 1, want_per_schedule=True)
thr = float(per[:, 0].median().item())
best, _ = eng.score_schedules(st, tm, 999.995, float(tm[-1]), P0, threshold=thr)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.reps):
    eng.score_schedules(st, tm, 999.995, float(tm[-1]), P0, threshold=thr, device_result=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.reps
print(json.dumps({"L": a.L, "schedules": 1 << a.L, "ms": dt * 1e3, "schedules_per_s": (1 << a.L) / dt, "threshold": thr,
                  **{k: best[k] for k in ("argmin_sum_index", "first_feasible_index", "n_feasible", "n_near_threshold")}}))
