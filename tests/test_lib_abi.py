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
    for name in ("CV6", "CA8", "BIAS9"):
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
