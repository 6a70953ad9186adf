"""How fast does cudaHostRegister pin a caller's pageable numpy array in place (the alternative to staging it
through pinned chunk buffers)?  python tools/host_register_probe.py [--gib 4]"""
import argparse
import json
import time

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gib", type=float, default=4.0)
a = ap.parse_args()
torch.cuda.init()
rt = torch.cuda.cudart()
n = int(a.gib * (1 << 30))
out = {"gib": a.gib, "thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()}
arr = np.empty(n, dtype=np.uint8)
arr[::4096] = 1  # touch every page
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk_mib in (0, 128):
    res = []
    for rep in range(3):
        if chunk_mib == 0:
            t0 = time.perf_counter()
            rc = rt.cudaHostRegister(arr.ctypes.data, n, 0)
            t1 = time.perf_counter()
            rt.cudaHostUnregister(arr.ctypes.data)
            t2 = time.perf_counter()
            res.append({"register_ms": (t1 - t0) * 1e3, "unregister_ms": (t2 - t1) * 1e3, "rc": int(rc)})
        else:
            c = chunk_mib << 20
            t0 = time.perf_counter()
            for off in range(0, n, c):
                rt.cudaHostRegister(arr.ctypes.data + off, min(c, n - off), 0)
            t1 = time.perf_counter()
            for off in range(0, n, c):
                rt.cudaHostUnregister(arr.ctypes.data + off)
            t2 = time.perf_counter()
            res.append({"register_ms": (t1 - t0) * 1e3, "unregister_ms": (t2 - t1) * 1e3})
    out[f"chunk_{chunk_mib}_mib"] = res
print(json.dumps(out))
