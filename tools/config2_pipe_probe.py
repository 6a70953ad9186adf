"""Where does a chunk of the pipelined rate sweep (BASELINE configs[1]) spend its time?  For a bounded stream:
the pipelined job, the selection walk alone, and the gain-applying launches of every rate alone and together
(selection precomputed).  python tools/config2_pipe_probe.py [--seconds 420] [--seeds 4096]"""
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
ap.add_argument("--seconds", type=float, default=420.0)
ap.add_argument("--chunk", type=int, default=25200)
a = ap.parse_args()
rates = list(range(10, 121, 10))
ev = synth.make_indexed_sensor_data(int(a.seconds * 210) + 3, seed=5)[3:]
ps = streams.PackedStream.from_events(ev)
eng = Engine("CA15")
dev = eng.device
base = torch.from_numpy(ps.data[1:]).to(dev)
E = int(base.shape[0])
gen = torch.Generator(device=dev).manual_seed(0)
N = a.seeds
bufs = [base[:a.chunk + 512, :, None] + 0.05 * torch.randn((a.chunk + 512, 9, N), generator=gen, device=dev, dtype=torch.float64)
        for _ in range(2)]
cnt = {"c": 0}


def recycled(e0, e1):
    b = bufs[cnt["c"] % 2]
    cnt["c"] += 1
    return b[:e1 - e0]


x0 = np.zeros((N, 15))
P0 = eng.spec.default_P0()
t_prev = float(ps.times[0])


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


res = {"events": E, "seeds": N, "chunks": E // a.chunk + 1, "order": os.environ.get("B200KF_SWEEP_ORDER", "descending")}
ms, out = timed(lambda: eng.scheduled_sweep_pipelined(ps.types[1:], ps.times[1:], rates, x0, P0, t_prev, recycled, a.chunk,
                                                       payload_depth=int(os.environ.get("PIPE_DEPTH", "2"))))
steps = float(sum(out[f]["n_sel"] for f in rates)) * N
res["pipelined_ms"] = ms
res["pipelined_gbs"] = steps * 72 / (ms * 1e-3) / 1e9
# selection alone (with and without gain records)
types_d, times_d = torch.from_numpy(ps.types[1:]).to(dev), torch.from_numpy(ps.times[1:]).to(dev)
P0r = np.repeat(P0[None], len(rates), axis=0)
for wg in (False, True):
    ms, got = timed(lambda: eng.select_greedy(types_d, times_d, [1.0 / f for f in rates], [t_prev] * len(rates), P0r, want_gains=wg))
    res["selection_ms_gains" if wg else "selection_ms"] = ms
sels, Pf, recs = got
# gain-applying launches, selection precomputed: every rate alone over ONE chunk of events, then all together
e1 = min(E, a.chunk)
pay = bufs[0][:e1]
x_d = torch.zeros((N, 15), dtype=torch.float64, device=dev)
alone = {}
engines = [eng] + [Engine("CA15") for _ in rates[1:]]
strs = [torch.cuda.Stream(device=dev) for _ in rates]
rows_all = []
for i, f in enumerate(rates):
    k = int(np.searchsorted(sels[i], e1))
    rows_all.append((k, torch.from_numpy(sels[i][:k].astype(np.int32)).to(dev)))
for i, f in enumerate(rates):
    k, rows = rows_all[i]
    ms, _ = timed(lambda: engines[i].run_gains(recs[i, :k], pay, x_d, step_rows=rows, inplace=True))
    alone[f] = {"steps": k, "ms": ms, "gbs": k * N * 72 / (ms * 1e-3) / 1e9}
res["one_chunk_each_rate_alone"] = alone


def together(order, small="none"):
    cur = torch.cuda.current_stream(dev)
    longest = max(k for k, _ in rows_all)
    for i in order:
        k, rows = rows_all[i]
        strs[i].wait_stream(cur)
        with torch.cuda.stream(strs[i]):
            st = small == "all" or (small == "heavy" and 2 * k > longest)
            engines[i].run_gains(recs[i, :k], pay, x_d.clone(), step_rows=rows, stream=strs[i], inplace=True, small_tiles=st)
    for s in strs:
        cur.wait_stream(s)


tot_steps = sum(k for k, _ in rows_all) * N
for name, order in (("descending", sorted(range(len(rates)), key=lambda j: -rates[j])),
                    ("ascending", sorted(range(len(rates)), key=lambda j: rates[j])),
                    ("first_wave_ascending", [rates.index(v) for v in (50, 60, 70, 80, 90, 100, 110, 120, 40, 30, 20, 10)])):
    ms, _ = timed(lambda: together(order), reps=3)
    res["one_chunk_together_" + name] = {"ms": ms, "gbs": tot_steps * 72 / (ms * 1e-3) / 1e9}
desc = sorted(range(len(rates)), key=lambda j: -rates[j])
for small in ("heavy", "all"):
    ms, _ = timed(lambda: together(desc, small), reps=3)
    res["one_chunk_together_descending_small_tiles_" + small] = {"ms": ms, "gbs": tot_steps * 72 / (ms * 1e-3) / 1e9}
k, rows = rows_all[-1]
ms, _ = timed(lambda: engines[-1].run_gains(recs[len(rates) - 1, :k], pay, x_d, step_rows=rows, inplace=True, small_tiles=True))
res["one_chunk_120hz_alone_small_tiles"] = {"steps": k, "ms": ms, "gbs": k * N * 72 / (ms * 1e-3) / 1e9}
print(json.dumps(res))
