"""HBM ceilings of this GPU for the read / write mixes the filter kernels produce (torch element-wise kernels,
1 Gi float64 = 8.6 GB per buffer, CUDA events): pure read (sum), pure write (fill), copy (1 read : 1 write, the
MEASURED_PEAKS.json method) and 3 reads : 7 writes (trajectory mode: 24 B in, 56 B out per filter-step).
Prints one JSON line."""
import json

import torch

n = 1 << 27  # 128 Mi float64 = 1 GiB per buffer
dev = torch.device("cuda", 0)
a = torch.ones(n * 3, dtype=torch.float64, device=dev)
outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(7)]


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
ms = timed(lambda: a.sum())
res["read_gbs"] = 3 * n * 8 / ms / 1e6
ms = timed(lambda: [o.fill_(1.5) for o in outs])
res["write_gbs"] = 7 * n * 8 / ms / 1e6
ms = timed(lambda: outs[0].copy_(outs[1]))
res["copy_gbs"] = 2 * n * 8 / ms / 1e6
x0, x1, x2 = a[:n], a[n:2 * n], a[2 * n:]


def mix():  # 3 reads, 7 writes per element, each output one pass (torch fuses nothing: 7 kernels re-reading is avoided
    # by writing outputs from a single read each; reads 3 + 4 re-reads of cached? no: count the actual traffic)
    torch.add(x0, x1, out=outs[0])      # 2 r 1 w
    torch.mul(x2, 2.0, out=outs[1])     # 1 r 1 w
    for o in outs[2:]:
        o.fill_(0.5)                    # 5 w
ms = timed(mix)
res["mix_3r_7w_gbs"] = 10 * n * 8 / ms / 1e6
print(json.dumps(res))
