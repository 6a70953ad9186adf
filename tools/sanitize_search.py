"""Small schedule searches over every path of the final search kernels, written for `compute-sanitizer --tool
memcheck / racecheck` (the shared-memory DFS levels): prefix-sharing kernel at 3, 4 and 5 levels, strided slices,
ragged ranges (flat head / tail), per-schedule rows, every model.  compute-sanitizer is closed on the pool this
round ran on (tools/gpu_round2_zzf.sh got exit code 86), so the script was run plainly; the bit-for-bit parity
tests in tests/test_gpu_schedules.py are the check.  python tools/sanitize_search.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sensorfusion_kalmanfilter_b200.engine import Engine  # noqa: E402

rng = np.random.default_rng(0)
for name in ("CA15", "CV6", "BIAS9", "CA8"):
    e = Engine(name)
    for L in ((15, 17, 20, 22) if name == "CA15" else (16,)):
        st = np.where(np.arange(L) % 4 == 1, 0, 1).astype(np.uint8)
        tm = 10.0 + np.cumsum(0.005 + 0.001 * rng.random(L))
        P0 = e.spec.default_P0() * 1e-2
        args = (st, tm, 9.99, float(tm[-1]), P0)
        full, _ = e.score_schedules(*args, threshold=-60.0)
        rag, per = e.score_schedules(*args, first=3, count=(1 << L) - 7, threshold=-60.0, want_per_schedule=(L <= 20))
        parts = [e.score_schedules(*args, threshold=-60.0, low_bits=r, n_low=2)[0] for r in range(4)]
        sl, per2 = e.score_schedules(*args, threshold=-60.0, low_bits=1, n_low=1, first=5, count=(1 << (L - 1)) - 9,
                                     want_per_schedule=(L <= 20))
        print(name, L, full["argmin_sum_index"], rag["n_feasible"], sum(p["n_feasible"] for p in parts), sl["n_feasible"])
        assert sum(p["n_feasible"] for p in parts) == full["n_feasible"]
