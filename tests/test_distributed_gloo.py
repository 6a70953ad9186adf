"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding of the schedule space (contiguous ranges,
slices by the earliest decisions, small windows replicated) and the deterministic all-gather + reduce of per-rank
search results."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from sensorfusion_kalmanfilter_b200.distributed import merge_best, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_best(rank, first, count):
    """A deterministic per-shard result: argmin score ties across ranks, feasibility only on rank 1."""
    none = (1 << 64) - 1
    return {"argmin_sum_score": -12.5, "argmin_sum_index": first + 3, "first_feasible_key": none if rank == 0 else 77,
            "first_feasible_index": None if rank == 0 else first + 5, "n_feasible": rank * 4,
            "n_near_threshold": 1 - rank}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from sensorfusion_kalmanfilter_b200.distributed import allgather_best
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(1 << 10, world, rank)
    out = allgather_best(_fake_best(rank, first, count))
    # the device-resident form of the same reduce (one all_gather_into_tensor of the raw 48-byte b200kf_best)
    import ctypes
    import torch
    from sensorfusion_kalmanfilter_b200 import _lib
    from sensorfusion_kalmanfilter_b200.distributed import allgather_best_device
    fb = _fake_best(rank, first, count)
    none = (1 << 64) - 1
    raw = _lib.Best(fb["argmin_sum_score"], fb["argmin_sum_index"], fb["first_feasible_key"],
                    none if fb["first_feasible_index"] is None else fb["first_feasible_index"], fb["n_feasible"],
                    fb["n_near_threshold"])
    best_d = torch.frombuffer(bytearray(bytes(raw)), dtype=torch.uint8).clone()
    assert best_d.numel() == ctypes.sizeof(_lib.Best)
    out2 = allgather_best_device(best_d)
    assert out2 == out, (out2, out)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allgather_best_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    parts = [_fake_best(r, *shard_range(1 << 10, world, r)) for r in range(world)]
    want = merge_best(parts)
    assert got[0] == want and got[1] == want  # every rank holds the same reduced answer
    assert want["argmin_sum_index"] == 3  # tie on the score -> lowest index (rank 0's shard)
    assert want["first_feasible_index"] == 512 + 5 and want["n_feasible"] == 4


class _FakeEngine:
    """Stands in for Engine.score_schedules on a CPU box: scores a toy objective over the slice it is asked for
    (sum score = popcount-weighted value with a unique minimum, feasible iff the schedule index is a multiple of 5)
    and records how it was called."""

    def __init__(self):
        self.calls = []

    def score_schedules(self, types, times, t_prev, t_end, P0, *, first=0, count=None, threshold=0.0,
                        device_result=False, low_bits=0, n_low=0, **kw):
        import torch
        from sensorfusion_kalmanfilter_b200 import _lib
        from sensorfusion_kalmanfilter_b200.kf_workers import _lex_key
        L = len(types)
        if count is None:
            count = (1 << (L - n_low)) - first
        self.calls.append({"first": first, "count": count, "low_bits": low_bits, "n_low": n_low, "device": device_result})
        none = (1 << 64) - 1
        sched = [low_bits | (j << n_low) for j in range(first, first + count)]
        sched = [s for s in sched if s != 0]
        score = {s: float((s * 2654435761) % 1000003) for s in sched}
        arg = min(sched, key=lambda s: (score[s], s))
        feas = [s for s in sched if s % 5 == 0]
        key = min((_lex_key(s, L) for s in feas), default=none)
        ffi = min(feas, key=lambda s: _lex_key(s, L)) if feas else none
        raw = _lib.Best(score[arg], arg, key, ffi, len(feas), 0)
        if device_result:
            return torch.frombuffer(bytearray(bytes(raw)), dtype=torch.uint8).clone(), None
        from sensorfusion_kalmanfilter_b200.engine import Engine
        return Engine.best_from_bytes(bytes(raw)), None


def _search_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    from sensorfusion_kalmanfilter_b200.distributed import search_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for L, limit in ((12, 8), (12, 22)):  # sharded by the earliest decision / replicated on every rank
        eng = _FakeEngine()
        types, times = np.zeros(L, np.uint8), np.arange(L, dtype=np.float64)
        out[(L, limit)] = (search_sharded(eng, types, times, 0.0, float(L), np.eye(2), 0.0, replicate_up_to_log2=limit),
                           eng.calls)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_search_sharded_slices_by_the_earliest_decision_and_replicates_small_windows():
    """search_sharded on two ranks with a stand-in engine: above the replication limit rank r scores the schedules
    whose first decision is r (one strided call over 2^(L-1) indices, result kept in device form for the all-gather);
    at or below it every rank scores the whole window and no collective runs.  Either way both ranks end with the
    answer of the unsharded search."""
    import numpy as np
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_search_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    L = 12
    whole, _ = _FakeEngine().score_schedules(np.zeros(L, np.uint8), None, 0, 0, None)
    for r in range(world):
        best, calls = got[r][(L, 8)]
        assert calls == [{"first": 0, "count": 1 << (L - 1), "low_bits": r, "n_low": 1, "device": True}]
        for k in ("argmin_sum_score", "argmin_sum_index", "first_feasible_index", "n_feasible"):
            assert best[k] == whole[k], (r, k)
        best, calls = got[r][(L, 22)]
        assert calls == [{"first": 0, "count": 1 << L, "low_bits": 0, "n_low": 0, "device": False}]
        assert best == whole
