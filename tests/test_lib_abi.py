"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200kf.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from sensorfusion_kalmanfilter_b200 import _lib, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200kf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200kf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200kf.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes prototypes out of sync with the header"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.ModelDesc) == 8 + 8 * (16 + 4 + 16 + 1)
    assert ctypes.sizeof(_lib.Best) == 48


def test_version_and_default_desc():
    assert _lib.version() == (0, 1)
    lib = _lib.load()
    d = _lib.ModelDesc()
    _lib.check(lib.b200kf_model_default_desc(_lib.MODEL_CA15, _lib.F64, ctypes.byref(d)))
    spec = models.default_spec("CA15")
    np.testing.assert_array_equal(np.array(d.q[:15]), spec.q)
    np.testing.assert_array_equal(np.array(d.r_imu[:15]), spec.r_imu)
    np.testing.assert_array_equal(np.array(d.r_gps[:3]), spec.r_gps)
    assert d.accel_coef == 0.5
    for name in ("CV6", "CA8", "BIAS9", "CVBU6"):
        _lib.check(lib.b200kf_model_default_desc(models.MODEL_IDS[name], _lib.F64, ctypes.byref(d)))
        s = models.default_spec(name)
        np.testing.assert_array_equal(np.array(d.q[:s.n]), s.q)
        np.testing.assert_array_equal(np.array(d.r_imu[:s.n]), s.r_imu)
    assert lib.b200kf_model_default_desc(17, _lib.F64, ctypes.byref(d)) == _lib.ERR_UNSUPPORTED
    assert "unknown model" in _lib.last_error()


def test_no_cpu_fallback():
    """Without a GPU the product path must refuse, not compute on the CPU."""
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    d = models.default_spec("CV6").to_desc()
    h = ctypes.c_void_p(0)
    assert lib.b200kf_model_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in _lib.last_error()
    from sensorfusion_kalmanfilter_b200.engine import Engine
    with pytest.raises(_lib.B200KFError):
        Engine("CV6")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sensorfusion-kalmanfilter_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_struct_layout_best_product():
    assert ctypes.sizeof(_lib.BestProduct) == 32


def test_header_is_plain_c_and_links_without_python(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (no C++, no torch, no Python) includes
    include/b200kf.h, links libb200kf.so and calls it -- what a cgo / JNI / N-API binding would do."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "b200kf.h"
int main(void) {
  int major = -1, minor = -1, count = -1;
  b200kf_model_desc d;
  b200kf_best best;
  b200kf_best_product bp;
  char err[256];
  if (b200kf_version(&major, &minor) != B200KF_OK) return 2;
  if (b200kf_model_default_desc(B200KF_MODEL_CA15, B200KF_F64, &d) != B200KF_OK) return 3;
  if (b200kf_model_default_desc(99, B200KF_F64, &d) != B200KF_ERR_UNSUPPORTED) return 4;
  b200kf_last_error(err, sizeof err);
  if (strstr(err, "unknown model") == NULL) return 5;
  b200kf_device_count(&count);
  printf("%d.%d devices=%d q0=%g r_gps0=%g sizeof(best)=%zu sizeof(best_product)=%zu\n", major, minor, count,
         d.q[0], d.r_gps[0], sizeof best, sizeof bp);
  return 0;
}
''')
    exe = tmp_path / "abi"
    pkg = os.path.join(ROOT, "sensorfusion-kalmanfilter_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", pkg, "-lb200kf", "-Wl,-rpath," + pkg]
    cuda_lib = "/usr/local/cuda/lib64"
    if os.path.isdir(cuda_lib):
        cmd += ["-Wl,-rpath-link," + cuda_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=cuda_lib + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("0.1 devices=") and "q0=5 r_gps0=3 sizeof(best)=48 sizeof(best_product)=32" in r.stdout
