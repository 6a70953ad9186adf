"""Builds libb200kf.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python sensorfusion-kalmanfilter_b200/csrc/build.py [--force]

One translation unit per (kernel family, model id) so the four models compile in parallel.
Objects go to csrc/build/ (git-ignored); the shared library lands next to the Python package so
it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
BUILD = os.path.join(HERE, "build")
# B200KF_VARIANT=<tag> + B200KF_EXTRA_FLAGS="-D..." build an experimental libb200kf_<tag>.so beside the
# default library (select it at run time with B200KF_LIB); used for A/B timing on the GPU box.
VARIANT = os.environ.get("B200KF_VARIANT", "")
EXTRA = os.environ.get("B200KF_EXTRA_FLAGS", "").split()
LIB = os.path.join(PKG, f"libb200kf_{VARIANT}.so" if VARIANT else "libb200kf.so")
if VARIANT:
    BUILD = os.path.join(HERE, "build_" + VARIANT)
NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-std=c++17", "-O3", "--expt-relaxed-constexpr", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xptxas", "-v"] + ARCH + EXTRA
MODELS = (0, 1, 2, 3, 4)
HEADERS = ["kf_core.cuh", "kf_axis.cuh", "kf_internal.h", os.path.join(ROOT, "include", "b200kf.h")]


def _units():
    units = [("kf_api", "kf_api.cu", [])]
    for mid in MODELS:
        units.append((f"kf_run_model_{mid}", "kf_run_model.cu", [f"-DB200KF_INST_MODEL={mid}"]))
        units.append((f"kf_sched_model_{mid}", "kf_sched_model.cu", [f"-DB200KF_INST_MODEL={mid}"]))
    return units


def _digest(src, extra):
    h = hashlib.sha256()
    for f in [src] + HEADERS:
        with open(os.path.join(HERE, f) if not os.path.isabs(f) else f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(CFLAGS + extra).encode())
    return h.hexdigest()


def _compile(unit, force):
    name, src, extra = unit
    obj = os.path.join(BUILD, name + ".o")
    stamp = obj + ".sha"
    dig = _digest(src, extra)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return name, obj, None
    cmd = [NVCC] + CFLAGS + extra + ["-c", os.path.join(HERE, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    with open(os.path.join(BUILD, name + ".ptxas.log"), "w") as fh:
        fh.write(r.stderr)
    return name, obj, r.stderr


def build(force=False, verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    units = _units()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        results = list(ex.map(lambda u: _compile(u, force), units))
    objs = [o for _, o, _ in results]
    rebuilt = [n for n, _, log in results if log is not None]
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"libb200kf.so: rebuilt {len(rebuilt)} of {len(units)} objects -> {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
