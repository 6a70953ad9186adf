"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference modules.

Imports /root/reference/kf_workers.py and /root/reference/hw5_2.py with empty
stand-ins for the third-party modules the reference imports but the Kalman
arithmetic never touches (`utm`, `matplotlib*`, `mpl_toolkits*`;
kf_workers.py:4-9, hw5_2.py:3-5).  Nothing is copied: the modules are imported
from where they lie.

/root/reference exists only in the build container, never on the GPU box, so
this loader is used by exactly two things: `oracle/make_golden.py` (which
writes tests/golden/*.npz) and the optional `-m "not gpu"` differential tests,
which skip when the directory is absent.  The product package must never
import this file.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("B200KF_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "kf_workers.py"))


def _stub(name: str) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    return mod


def load_reference():
    """Return (kf_workers, hw5_2) imported unmodified from REFERENCE_DIR."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_DIR}")
    for name in ("utm", "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    m3d = sys.modules["mpl_toolkits.mplot3d"]
    if not hasattr(m3d, "Axes3D"):
        m3d.Axes3D = object
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "animation"):
        if not hasattr(mpl, sub):
            setattr(mpl, sub, sys.modules["matplotlib." + sub])
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import kf_workers as ref_kf_workers  # noqa: E402
    import hw5_2 as ref_hw5_2  # noqa: E402
    return ref_kf_workers, ref_hw5_2


def load_notebook_class():
    """KF_SensorFusion (and Scheduler) as defined in cells 0-2 of the reference's
    KF_SensorFusion.ipynb, exec'd unmodified -- the only place run_brute_force_kalman_filter
    (product-of-queues search, RMSE objective; KF_SensorFusion.ipynb:804-954) exists."""
    import json
    load_reference()  # installs the utm / matplotlib stand-ins
    with open(os.path.join(REFERENCE_DIR, "KF_SensorFusion.ipynb")) as f:
        nb = json.load(f)
    ns = {"__name__": "kf_notebook"}
    code_cells = [c for c in nb["cells"] if c["cell_type"] == "code"]
    for c in code_cells[:3]:
        # IPython shell magics ("!pip install ...") are not Python: dropped, nothing else is touched
        src = "".join(ln for ln in c["source"] if not ln.lstrip().startswith(("!", "%")))
        exec(compile(src, "KF_SensorFusion.ipynb", "exec"), ns)
    return ns["KF_SensorFusion"]
