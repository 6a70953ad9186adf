"""Throughput of the filter-loop kernels for every model / precision / mode (diagnostic; results are
copied into profiles/).  python tools/bench_models.py [--filters N] [--tsteps T]"""
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
ap.add_argument("--tsteps", type=int, default=500)
ap.add_argument("--models", default="CV6,CA8,BIAS9,CA15")
a = ap.parse_args()
dev = torch.device("cuda", 0)
rows = []
for name in a.models.split(","):
    for prec in ("f64", "f32"):
        eng = Engine(name, precision=prec)
        sp = eng.spec
        N, T = a.filters, a.tsteps
        types, dts = synth.gps_imu_schedule(T)
        types_d, dts_d = torch.from_numpy(types).to(dev), torch.from_numpy(dts).to(dev)
        pay = fill_payload(eng, types_d, N, seed=7)
        for mode in ("auto", "decoupled", "dense"):
            for full in (False, True):
                x = torch.zeros((N, sp.n), dtype=eng.real, device=dev)
                P = torch.from_numpy(np.repeat(sp.default_P0()[None], N, axis=0)).to(device=dev, dtype=eng.real).contiguous()
                kw = dict(mode=mode, inplace=True)
                if full:
                    kw.update(want_agg=False, traj_ncomp=min(6, sp.n), want_logdet=True)

                def run():
                    o = eng.run_batch(types_d, dts_d, pay, x, P, **kw)
                    del o
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                reps = 3
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                rb = 4 if prec == "f32" else 8
                bytes_step = sp.d * rb + ((min(6, sp.n) + 1) * rb if full else 0)
                row = {"model": name, "precision": prec, "mode": mode, "outputs": "trajectory" if full else "aggregate",
                       "filters": N, "steps": T, "ms": round(ms, 3), "filter_steps_per_s": N * T / (ms * 1e-3),
                       "alg_bytes_per_step": bytes_step, "gbs": N * T * bytes_step / (ms * 1e-3) / 1e9}
                rows.append(row)
                print(json.dumps(row), flush=True)
        del pay
        eng.close()
