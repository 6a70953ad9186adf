"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding of the schedule
index space and the deterministic all-gather + reduce of per-rank search results."""
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
