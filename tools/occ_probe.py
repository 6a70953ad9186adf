"""How many CTAs of a small shared-covariance launch share an SM?  Times k concurrent 19-CTA launches
(4096 CA15 filters each, one stream each) against one: with two resident CTAs per SM, 12 launches (228 CTAs)
take about as long as one.   python tools/occ_probe.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200 import synth  # noqa: E402
from sensorfusion_kalmanfilter_b200.engine import Engine, fill_payload  # noqa: E402

dev = torch.device("cuda", 0)
N, T = 4096, 6000
engs = [Engine("CA15") for _ in range(12)]
sp = engs[0].spec
types, dts = synth.gps_imu_schedule(T)
types_d, dts_d = torch.from_numpy(types).to(dev), torch.from_numpy(dts).to(dev)
pay = fill_payload(engs[0], types_d, N, seed=3)
x = [torch.zeros((N, sp.n), dtype=torch.float64, device=dev) for _ in engs]
P = [torch.from_numpy(np.repeat(sp.default_P0()[None], N, axis=0)).to(dev).contiguous() for _ in engs]
streams = [torch.cuda.Stream(device=dev) for _ in engs]
out = {}
for k in (1, 4, 8, 12):
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            streams[i].wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(streams[i]):
                engs[i].run_batch(types_d, dts_d, pay, x[i], P[i], inplace=True, stream=streams[i], mode="shared")
        for i in range(k):
            torch.cuda.current_stream(dev).wait_stream(streams[i])
        e1.record()
        torch.cuda.synchronize()
    out[k] = round(e0.elapsed_time(e1), 3)
print(json.dumps({"launch": "CA15 shared covariance, 4096 filters x 6000 steps (19 CTAs of 224 filters)",
                  "ms_for_k_concurrent_launches": out}))
