import os
import subprocess
import sys

import pytest

# several contexts of one process stand in for several GPUs in tests/test_multigpu_gpu.py: their barrier kernels spin on one
# another, so every stream needs a hardware queue of its own (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the checker (oracle/) and the product library exist; both are in-tree builds."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all", "-j8"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "harc_b200", "libharcgpu.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def workroot(tmp_path_factory):
    return str(tmp_path_factory.mktemp("harc"))
