"""BASELINE configs[1] at a bounded scale: sampling-rate sweep 10..120 Hz (the 12 rates of the
reference's sampling_sweep/) x Monte-Carlo noise seeds of the 15-state filter with greedy
one-event-per-window scheduling (run_kalman_filter_scheduled, kf_workers.py:826-957).

    python tools/config2_sweep.py [--seeds 4096] [--seconds 60] [--no-concurrent]

The base stream is synthetic (200 Hz IMU / 10 Hz GPS, reference layout); every seed adds its own
measurement noise.  One covariance-only selection pass for all rates, then per rate one batch of
`seeds` filters sharing the selected schedule and the covariance.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import streams, synth  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=4096)
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--no-concurrent", action="store_true")
ap.add_argument("--use-gains", action="store_true", help="filters apply the gain records of the selection pass")
a = ap.parse_args()
rates = list(range(10, 121, 10))
n_events = int(a.seconds * 210)
ev = synth.make_indexed_sensor_data(n_events + 3, seed=5)[3:]
assert ev[0][1] == "GPS"
ps = streams.PackedStream.from_events(ev)
dev = torch.device("cuda", 0)
eng = Engine("CA15")
g = torch.Generator(device=dev).manual_seed(0)
base = torch.from_numpy(ps.data[1:]).to(dev)                                   # [E, 9]
E = base.shape[0]
pay = base[:, :, None] + 0.05 * torch.randn((E, 9, a.seeds), generator=g, device=dev, dtype=torch.float64)
x0 = np.zeros((a.seeds, 15))
x0[:, :3] = ps.data[0, :3]
P0 = eng.spec.default_P0()
kw = dict(concurrent=not a.no_concurrent, use_gains=a.use_gains)
for _ in range(2):
    out = eng.scheduled_sweep(ps.types[1:], ps.times[1:], pay, rates, x0, P0, float(ps.times[0]), **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    out = eng.scheduled_sweep(ps.types[1:], ps.times[1:], pay, rates, x0, P0, float(ps.times[0]), **kw)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
steps = sum(len(out[f]["sel"]) for f in rates) * a.seeds
print(json.dumps({"config": "rate sweep 10-120 Hz x %d seeds, CA15, greedy scheduling, %.0f s stream (%d events)"
                  % (a.seeds, a.seconds, E), "concurrent_streams": not a.no_concurrent, "use_gains": a.use_gains, "filters": a.seeds * len(rates),
                  "filter_steps": steps, "wall_s": dt, "filter_steps_per_s": steps / dt,
                  "selected_per_rate": {str(f): int(len(out[f]["sel"])) for f in rates}}))
