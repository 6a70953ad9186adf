"""Per-pass device time of the headline filter kernel over many back-to-back passes, next to the
clocks nvidia-smi reports: shows whether the step time drifts under sustained load.

    python tools/drift_probe.py [passes=60] [mode=symmetric] [reset]
"""
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import synth  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine, fill_payload  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 60
mode = sys.argv[2] if len(sys.argv) > 2 else "symmetric"
reset = len(sys.argv) > 3 and sys.argv[3] == "reset"
N, T = 131072, 10000
dev = torch.device("cuda", 0)
os.environ["B200KF_SHARED"] = "0" if mode != "shared" else "1"
eng = Engine("CV6")
sp = eng.spec
types, dts = synth.gps_imu_schedule(T)
types_d, dts_d = torch.from_numpy(types).to(dev), torch.from_numpy(dts).to(dev)
payload = fill_payload(eng, types_d, N, seed=1234)
scale = 1.0 + (np.arange(N) % 1024) / 1024.0 if mode != "shared" else np.ones(N)
P0 = torch.from_numpy(sp.default_P0()[None] * scale[:, None, None]).to(dev).contiguous()
x = torch.zeros((N, sp.n), dtype=torch.float64, device=dev)
P = P0.clone()
q = "timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active"
smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
time.sleep(0.3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(passes + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(passes):
    if reset:
        x.zero_()
        P.copy_(P0)
    eng.run_batch(types_d, dts_d, payload, x, P, mode=mode, inplace=True)
    ev[i + 1].record()
torch.cuda.synchronize()
time.sleep(0.1)
smi.terminate()
out, _ = smi.communicate()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(passes)]
print("per-pass ms:", " ".join(f"{m:.2f}" for m in ms))
print("x abs max %.3e  P abs max %.3e  P abs min diag %.3e" % (x.abs().max().item(), P.abs().max().item(),
                                                             torch.diagonal(P, dim1=1, dim2=2).min().item()))
rows = [ln for ln in out.strip().splitlines()]
print("nvidia-smi samples (every 5th):")
for ln in rows[::5]:
    print("  ", ln)
