"""Multi-GPU plumbing: one process per GPU, sharding by index range (filters) or by the earliest decisions
(schedules), one tiny final collective.

The path shards naturally: filters are independent, and so are candidate
schedules, so rank r owns a contiguous range of filters, or the schedules whose earliest decisions spell r
(shard_strided), and nothing is exchanged during compute.  The
reference's own "parallelism" is a `multiprocessing.Pool(30)` whose chunks are awaited one by one
(kf_workers.py:1333-1346).  The only collective here is the end-of-search all-gather of each
rank's `b200kf_best` (48 bytes) followed by the same deterministic reduce on every rank -- an
all-reduce(MIN) could not carry the index nor define the tie order.
"""
from __future__ import annotations

import numpy as np

_NONE = None


def shard_range(total: int, world_size: int, rank: int):
    """Contiguous [first, first+count) of `total` items for `rank`; sizes differ by at most 1."""
    base, rem = divmod(int(total), int(world_size))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def merge_best(parts):
    """Deterministic reduce of per-shard search results (dicts from Engine.score_schedules):
    min sum-score with lowest-index ties; min enumeration key for the first feasible schedule."""
    out = {"argmin_sum_score": float("inf"), "argmin_sum_index": None, "first_feasible_key": (1 << 64) - 1,
           "first_feasible_index": None, "n_feasible": 0, "n_near_threshold": 0}
    for b in parts:
        if b["argmin_sum_index"] is not None:
            s, i = b["argmin_sum_score"], b["argmin_sum_index"]
            if (out["argmin_sum_index"] is None or s < out["argmin_sum_score"]
                    or (s == out["argmin_sum_score"] and i < out["argmin_sum_index"])):
                out["argmin_sum_score"], out["argmin_sum_index"] = s, i
        if b["first_feasible_index"] is not None and b["first_feasible_key"] < out["first_feasible_key"]:
            out["first_feasible_key"] = b["first_feasible_key"]
            out["first_feasible_index"] = b["first_feasible_index"]
        out["n_feasible"] += b["n_feasible"]
        out["n_near_threshold"] += b["n_near_threshold"]
    return out


_FIELDS = ("argmin_sum_score", "argmin_sum_index", "first_feasible_key", "first_feasible_index",
           "n_feasible", "n_near_threshold")


def _pack(b):
    none = (1 << 64) - 1
    ints = np.array([none if b["argmin_sum_index"] is None else b["argmin_sum_index"],
                     b["first_feasible_key"],
                     none if b["first_feasible_index"] is None else b["first_feasible_index"],
                     b["n_feasible"], b["n_near_threshold"]], dtype=np.uint64)
    return np.concatenate([np.array([b["argmin_sum_score"]], dtype=np.float64).view(np.uint64), ints])


def _unpack(row):
    none = (1 << 64) - 1
    row = np.asarray(row, dtype=np.uint64)
    score = float(row[:1].view(np.float64)[0])
    idx, key, ffi, nf, nn = (int(v) for v in row[1:])
    return {"argmin_sum_score": score, "argmin_sum_index": None if idx == none else idx,
            "first_feasible_key": key, "first_feasible_index": None if key == none else ffi,
            "n_feasible": nf, "n_near_threshold": nn}


def allgather_best(best, group=None, device=None):
    """All-gather every rank's result (6 x 8 bytes) and reduce identically on all ranks.
    Works on NCCL (device tensors) and gloo (CPU tensors, used by the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_best([best])
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cpu")
    if backend == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    mine = torch.from_numpy(_pack(best).view(np.int64).copy()).to(dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    parts = [_unpack(g.cpu().numpy().view(np.uint64)) for g in gathered]
    return merge_best(parts)


def allgather_best_device(best_d, group=None):
    """The same reduce with the host off the critical path: `best_d` is the raw 48-byte b200kf_best still on
    the device (Engine.score_schedules(device_result=True)); ONE all_gather_into_tensor of it, stream-ordered
    behind the scoring kernel (no host synchronisation in between), then ONE device->host copy of the
    world_size x 48 bytes and the deterministic merge (a 48-byte-per-rank reduce; the order of ties is part of
    the contract, so it is not an all-reduce)."""
    import torch
    import torch.distributed as dist
    from .engine import Engine
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return merge_best([Engine.best_from_bytes(best_d.cpu().numpy().tobytes())])
    if dist.get_backend(group) != "nccl":
        best_d = best_d.cpu()
    out = torch.empty((world * best_d.numel(),), dtype=best_d.dtype, device=best_d.device)
    dist.all_gather_into_tensor(out, best_d, group=group)
    raw = out.cpu().numpy().reshape(world, -1)
    return merge_best([Engine.best_from_bytes(raw[r].tobytes()) for r in range(world)])


def shard_strided(n_events: int, world_size: int, rank: int):
    """(low_bits, n_low) of rank `rank`'s slice of the 2^n_events subset schedules, or None.

    A schedule's cost grows with the events it uses, and the prefix-sharing kernel pays the events ABOVE an aligned
    index block once per schedule: with one contiguous range per rank, the rank whose top bits are all ones works
    1.46 x as long as rank 0 (8 ranks, 2^32 schedules: 23.9 vs 16.4 ms, tools/search_shard_probe.py), and finer
    contiguous blocks only fix more late events (8 complementary blocks per rank: 28.1 ms).  Fixing the EARLIEST
    log2(world) decisions instead costs nothing per schedule (they are filtered once per thread, before the shared
    prefix) and every rank walks the same tree over the remaining events.  None when the world size is not a
    power of two or the window is too short: contiguous ranges then."""
    world_size, rank = int(world_size), int(rank)
    if world_size <= 1 or (world_size & (world_size - 1)) != 0:
        return None
    n_low = world_size.bit_length() - 1
    if n_low > int(n_events):
        return None
    return rank, n_low


def search_sharded(engine, types, times, t_prev, t_end, P0, threshold, group=None, replicate_up_to_log2=22, **kw):
    """Brute-force search of all 2^L subset schedules sharded over the ranks of `group`: rank r takes the schedules
    whose earliest log2(world) decisions spell r (equal work, shard_strided; contiguous index ranges for other
    world sizes), no exchange during the search, one 48-byte all-gather at the end.

    A window of at most `replicate_up_to_log2` events is searched whole by EVERY rank instead (same kernel, same
    inputs, same answer on all of them, no collective): 2^22 subsets take one GPU 0.1 ms, less than an all-gather and
    its synchronisation cost (config 4's 2^20 on 8 GPUs: 0.2 ms sharded, 0.05 ms replicated)."""
    import torch.distributed as dist
    L = len(types)
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 1 and L <= int(replicate_up_to_log2):
        best, _ = engine.score_schedules(types, times, t_prev, t_end, P0, threshold=threshold, **kw)
        return merge_best([best])
    strided = shard_strided(L, world, rank)
    if strided is not None:
        kw = dict(kw, low_bits=strided[0], n_low=strided[1], first=0, count=None)
    else:
        first, count = shard_range(1 << L, world, rank)
        kw = dict(kw, first=first, count=count)
    best_d, _ = engine.score_schedules(types, times, t_prev, t_end, P0, threshold=threshold, device_result=True, **kw)
    return allgather_best_device(best_d, group=group)
