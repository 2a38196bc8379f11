import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree; build it (nvcc cross-compiles on CPU) if it is missing or stale."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ub_build", os.path.join(ROOT, "unet-lane-detection_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
