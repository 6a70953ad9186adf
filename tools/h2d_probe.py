"""Diagnostic: pinned H2D bandwidth and run_batch_host chunk-size sweep (not part of the product)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import synth
from sensorfusion_kalmanfilter_b200.engine import Engine, fill_payload
dev = torch.device("cuda", 0)
for mb in (64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"H2D pinned {mb} MiB: {mb / 1024 / dt:.1f} GiB/s")
    for _ in range(2): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"D2H pinned {mb} MiB: {mb / 1024 / dt:.1f} GiB/s")
eng = Engine("CV6")
N, T = 131072, 1250
types, dts = synth.gps_imu_schedule(T)
pay = fill_payload(eng, torch.from_numpy(types).to(dev), N)
ph = torch.empty((T, 3, N), dtype=torch.float64, pin_memory=True); ph.copy_(pay); torch.cuda.synchronize()
x = np.zeros((N, 6)); P = np.repeat(eng.spec.default_P0()[None], N, axis=0)
for cs in (0, 16, 64, 256, 1250):
    eng.run_batch_host(types, dts, ph, x, P, chunk_steps=cs)
    t0 = time.perf_counter()
    for _ in range(3): eng.run_batch_host(types, dts, ph, x, P, chunk_steps=cs)
    dt = (time.perf_counter() - t0) / 3
    print(f"run_batch_host chunk_steps={cs}: {dt*1e3:.1f} ms  {N*T/dt:.3e} steps/s  {T*3*N*8/dt/2**30:.1f} GiB/s")
