"""Sequences scored per second by the product-of-queues search (b200kf_score_products; the notebook's
run_brute_force_kalman_filter, for which the reference reports 13.9 sequences/s).

    python tools/bench_products.py [--windows 10] [--queue 4]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=10)
ap.add_argument("--queue", type=int, default=4)
a = ap.parse_args()
rng = np.random.default_rng(0)
W, q = a.windows, a.queue
E = W * q
times = 1000.0 + np.sort(rng.uniform(0, 0.02 * W, E))
types = (rng.random(E) < 0.9).astype(np.uint8)
pay = rng.normal(0, 0.1, (E, 9))
ref = rng.normal(0, 1.0, (E, 3))
eng = Engine("CA15")
P0 = np.diag([1000.0] * 3 + [100.0] * 9 + [1000.0] * 3)
for _ in range(2):
    best, _ = eng.score_products([q] * W, types, times, pay, ref, np.zeros(15), P0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    best, _ = eng.score_products([q] * W, types, times, pay, ref, np.zeros(15), P0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"workload": f"CA15, {W} windows x {q} candidates = {q ** W} sequences, RMSE objective", "ms": ms,
                  "sequences_per_s": q ** W / (ms * 1e-3), "best_index": best["index"], "best_rmse": best["score"]}))
