import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # libb200kf.so is a build artefact (git-ignored): compile it when a fresh checkout has none
    lib = os.path.join(ROOT, "sensorfusion-kalmanfilter_b200", "libb200kf.so")
    if not os.path.exists(lib):
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "b200kf_build", os.path.join(ROOT, "sensorfusion-kalmanfilter_b200", "csrc", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(verbose=False)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load
