"""Sharded brute-force schedule search under torchrun (NCCL): every rank scores its slice of the 2^L subsets (the
earliest log2(world) decisions fixed; windows of <= 22 events are searched whole by every rank, no collective) and all
ranks end with the same reduced answer (distributed.search_sharded).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/search_multi_gpu.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200.distributed import search_sharded  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine("CA15")
for L in (20, 24, 26):
    rng = np.random.default_rng(0)
    types = np.where(np.arange(L) % 7 == 3, 0, 1).astype(np.uint8)
    times = 1000.0 + np.cumsum(0.005 + 0.0005 * rng.random(L))
    P0 = eng.spec.default_P0() * 1e-3
    full, _ = eng.score_schedules(types, times, 999.995, float(times[-1]), P0, threshold=0.0)
    thr = full["argmin_sum_score"] / (L + 2) + 1.0  # a threshold some schedules meet
    ref, _ = eng.score_schedules(types, times, 999.995, float(times[-1]), P0, threshold=thr)
    for _ in range(2):
        got = search_sharded(eng, types, times, 999.995, float(times[-1]), P0, thr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        got = search_sharded(eng, types, times, 999.995, float(times[-1]), P0, thr)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    ok = all(got[k] == ref[k] for k in ("argmin_sum_index", "first_feasible_index", "n_feasible"))
    if rank == 0:
        print(json.dumps({"L": L, "world": world, "ms_per_search_incl_allgather": dt * 1e3,
                          "schedules_per_s": (1 << L) / dt, "matches_single_gpu": ok,
                          "first_feasible_index": got["first_feasible_index"], "n_feasible": got["n_feasible"]}))
    assert ok, (got, ref)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
