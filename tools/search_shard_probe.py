"""How long does ONE rank's shard of a 2^L subset search take on its own?  (bench.py's search_2p32 leg at N GPUs
= this + one 48-byte all-gather.)  python tools/search_shard_probe.py [--L 32] [--world 8] [--contiguous]
Default: the slices search_sharded uses (the log2(world) earliest decisions fixed, distributed.shard_strided);
--contiguous: one index range per rank (what it used before: the late events fixed, paid by every schedule)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sensorfusion_kalmanfilter_b200.distributed import shard_range, shard_strided  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--contiguous", action="store_true")
a = ap.parse_args()
eng = Engine("CA15")
st, tm = bench.search_window(a.L)
P0 = bench.warm_cov(eng)
t_prev, t_end = 999.995, float(tm[-1])
_, per = eng.score_schedules(st, tm, t_prev, t_end, P0, first=1, count=(1 << 16) - 1, want_per_schedule=True)
thr = float(per[:, 0].median().item())
res = {"L": a.L, "world": a.world, "sharding": "contiguous" if a.contiguous else "strided", "ranks": {}}
for rank in sorted({0, min(1, a.world - 1), a.world // 2 if a.world > 1 else 0, a.world - 1}):
    kw = {}
    if a.contiguous or shard_strided(a.L, a.world, rank) is None:
        first, count = shard_range(1 << a.L, a.world, rank)
    else:
        low, n_low = shard_strided(a.L, a.world, rank)
        first, count, kw = 0, 1 << (a.L - n_low), {"low_bits": low, "n_low": n_low}
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.score_schedules(st, tm, t_prev, t_end, P0, first=first, count=count, threshold=thr, device_result=True, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res["ranks"][rank] = {"first": first, "count": count, "ms": ms, "schedules_per_s": count / (min(ms[1:]) * 1e-3)}
print(json.dumps(res))
