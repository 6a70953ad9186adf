"""Which resource bounds the host-buffer (e2e) path when several GPUs stream at once?  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_matrix_probe.py

One rank per GPU, 1 GiB pinned buffers, cudaMemcpyAsync through torch (the copy b200kf_run_batch_host issues).
Measures pinned H2D (and D2H) bandwidth (a) of every GPU alone, (b) of GPU 0 together with one other GPU,
(c) of the first 2 / 4 / 8 GPUs together, next to (d) the HOST's own memory bandwidth (multi-threaded memcpy on
all ranks at once: what the CPU side can read + write at all) and the PCI / NUMA placement of every GPU.
Rank 0 prints one JSON line."""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
GIB = 1 << 30
h = torch.empty(GIB, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(GIB, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def copy_gbs(direction, active, reps=6):
    """GB/s of this rank's copies while the ranks in `active` copy at the same time (0.0 if not active)."""
    barrier()
    gbs = 0.0
    if rank in active:
        src, dst = (h, d) if direction == "h2d" else (d, h)
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        gbs = reps * GIB / (time.perf_counter() - t0) / 1e9
    barrier()
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]
    return [gbs]


def host_memcpy_gbs(active):
    a = torch.empty(GIB // 2, dtype=torch.uint8)
    b = torch.empty(GIB // 2, dtype=torch.uint8)
    a.fill_(3)
    b.copy_(a)
    barrier()
    gbs = 0.0
    if rank in active:
        t0 = time.perf_counter()
        for _ in range(6):
            b.copy_(a)
        gbs = 6 * 2 * (GIB // 2) / (time.perf_counter() - t0) / 1e9  # bytes read + written
    barrier()
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]
    return [gbs]


res = {"world": world, "torch_cpu_threads": torch.get_num_threads(), "host_cores": len(os.sched_getaffinity(0))}
pr = torch.cuda.get_device_properties(local)
bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
try:
    numa = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
except Exception:
    numa = None
info = [None] * world
if world > 1:
    dist.all_gather_object(info, {"rank": rank, "pci": bdf, "numa_node": numa})
else:
    info = [{"rank": rank, "pci": bdf, "numa_node": numa}]
res["gpus"] = info
for direction in ("h2d", "d2h"):
    solo = [copy_gbs(direction, {r})[r] for r in range(world)]
    res[direction + "_alone_gbs"] = solo
    if world > 1:
        res[direction + "_gpu0_with"] = {str(o): copy_gbs(direction, {0, o}) for o in sorted({1, 2, 4, world - 1} & set(range(1, world)))}
        for k in (2, 4, 8):
            if k <= world:
                v = copy_gbs(direction, set(range(k)))
                res[f"{direction}_first{k}_together_gbs"] = v
                res[f"{direction}_first{k}_aggregate_gbs"] = sum(v)
res["host_memcpy_one_rank_gbs"] = host_memcpy_gbs({0})[0]
allm = host_memcpy_gbs(set(range(world)))
res["host_memcpy_all_ranks_gbs"] = allm
res["host_memcpy_all_ranks_aggregate_gbs"] = sum(allm)
if rank == 0:
    try:
        res["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        pass
    try:
        res["numa_nodes"] = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
        res["meminfo_head"] = open("/proc/meminfo").read().splitlines()[:3]
    except Exception:
        pass
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
