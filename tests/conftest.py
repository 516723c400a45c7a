import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # native pieces are built in-tree; rebuild here when sources are newer (nvcc cross-compiles without a GPU)
    if not os.environ.get("ORBX_SKIP_BUILD"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tc2li_slam_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liborb_oracle.so"], stdout=subprocess.DEVNULL)


def _has_gpu():
    try:
        import ctypes
        import tc2li_slam_b200 as T
        h = ctypes.c_void_p()
        rc = T.lib().orbx_batch_create(100, 1.2, 2, 20, 7, 0, 128, 128, 1, ctypes.byref(h))
        if rc == 0:
            T.lib().orbx_batch_destroy(h)
        return rc == 0
    except Exception:
        return False


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAS_GPU
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    if HAS_GPU is None:
        HAS_GPU = _has_gpu()
    if not HAS_GPU:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_cases():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "appendix_c.json")) as f:
        return json.load(f)
