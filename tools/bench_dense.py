"""Throughput of the dense-covariance path of the 15-state filter (arbitrary coupled initial_pt,
kf_workers.py:642-645): three lanes per filter (kf_run_axis_kernel) against the thread-per-filter dense kernel.
python tools/bench_dense.py [--filters N] [--tsteps T]   (diagnostic; results are copied into profiles/)"""
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
ap.add_argument("--filters", type=int, default=131072)
ap.add_argument("--tsteps", type=int, default=200)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precisions", default="f64,f32")
ap.add_argument("--axis", default="1,0")
a = ap.parse_args()
dev = torch.device("cuda", 0)
N, T = a.filters, a.tsteps
rng = np.random.default_rng(0)
A = rng.normal(size=(1024, 15, 15))
P0 = A @ np.swapaxes(A, 1, 2) + 5.0 * np.eye(15)  # dense SPD, every state coupled
P0 = np.tile(P0, (N // 1024 + 1, 1, 1))[:N]
for prec in a.precisions.split(","):
    for axis in a.axis.split(","):
        os.environ["B200KF_AXIS_DENSE"] = axis
        eng = Engine("CA15", precision=prec)
        sp = eng.spec
        types, dts = synth.gps_imu_schedule(T)
        types_d, dts_d = torch.from_numpy(types).to(dev), torch.from_numpy(dts).to(dev)
        pay = fill_payload(eng, types_d, N, seed=7)
        P0_d = torch.from_numpy(P0).to(device=dev, dtype=eng.real).contiguous()
        for full in (False, True):
            x = torch.zeros((N, sp.n), dtype=eng.real, device=dev)
            P = P0_d.clone()
            kw = dict(mode="auto", inplace=True)
            if full:
                kw.update(want_agg=False, traj_ncomp=6, want_logdet=True)

            def run():
                o = eng.run_batch(types_d, dts_d, pay, x, P, **kw)
                del o
            run()
            torch.cuda.synchronize()
            ms = []
            for _ in range(a.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            row = {"model": "CA15", "covariance": "dense (coupled P0)", "precision": prec,
                   "kernel": "three lanes per filter" if axis == "1" else "thread per filter",
                   "outputs": "trajectory" if full else "aggregate", "filters": N, "steps": T,
                   "ms_min": round(min(ms), 3), "ms_mean": round(sum(ms) / len(ms), 3),
                   "filter_steps_per_s": N * T / (sum(ms) / len(ms) * 1e-3)}
            print(json.dumps(row), flush=True)
        del pay, P0_d
        eng.close()
