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


def record_parity(name, payload, fname="r2_parity.json"):
    """Parity figures the tests measure (noise floors, per-tensor gradient errors) are written next to the other GPU-run
    artefacts (gpurun_out/, merged back by gpurun) so they can be committed under profiles/."""
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        path = os.path.join(out_dir, fname)
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = payload
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
