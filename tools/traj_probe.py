"""Trajectory-mode launches of the per-filter ring kernel for ncu / A-B timing:
    python tools/traj_probe.py [--model CV6] [--filters 131072] [--tsteps 1000] [--reps 20] [--precision f64]
Prints one JSON line (ms per launch, GB/s of the 24 B in + 56 B out per filter-step)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import synth  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine, fill_payload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="CV6")
ap.add_argument("--filters", type=int, default=131072)
ap.add_argument("--tsteps", type=int, default=1000)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--precision", default="f64")
ap.add_argument("--ncomp", type=int, default=6)
ap.add_argument("--mode", default="auto")
a = ap.parse_args()
os.environ["B200KF_SHARED"] = "0"
eng = Engine(a.model, precision=a.precision)
sp = eng.spec
N, T = a.filters, a.tsteps
types, dts = synth.gps_imu_schedule(T)
types_d = torch.from_numpy(types).to(eng.device)
dts_d = torch.from_numpy(dts).to(eng.device)
pay = fill_payload(eng, types_d, N, seed=7)
x = torch.zeros((N, sp.n), dtype=eng.real, device=eng.device)
scale = 1.0 + (np.arange(N) % 1024) / 1024.0
P = torch.from_numpy(sp.default_P0()[None] * scale[:, None, None]).to(device=eng.device, dtype=eng.real).contiguous()


def one():
    o = eng.run_batch(types_d, dts_d, pay, x, P, want_agg=False, traj_ncomp=a.ncomp, want_logdet=True, inplace=True,
                      mode=a.mode)
    del o


for _ in range(3):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    one()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
rb = 4 if a.precision == "f32" else 8
byts = float(N) * T * (sp.d + a.ncomp + 1) * rb
print(json.dumps({"model": a.model, "precision": a.precision, "filters": N, "tsteps": T, "ms": ms, "steps_per_s": N * T / ms * 1e3,
                  "gbs": byts / ms / 1e6, "bulk_out": os.environ.get("B200KF_BULK_OUT", "1")}))
