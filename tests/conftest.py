import os
import sys

import numpy as np
import pytest

os.environ.setdefault("OMP_WAIT_POLICY", "passive")
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CUDA library and the oracle are built in-tree (nvcc cross-compiles here; the prebuilt .so travels to the GPU box)."""
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build_cuda()
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        ge.build_oracle()
    yield


@pytest.fixture(scope="session")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback to test)")
    from mr_slam_b200 import _lib
    _lib.init(0)
    return _lib


@pytest.fixture(scope="session")
def gicp_golden():
    return np.load(os.path.join(GOLDEN, "gicp_fixture.npz"))


@pytest.fixture(scope="session")
def readme_golden():
    return np.load(os.path.join(GOLDEN, "gicp_readme_fixture.npz"))


def rel_l2(ref, x):
    ref, x = np.asarray(ref, np.float64), np.asarray(x, np.float64)
    return float(np.linalg.norm(ref - x) / (np.linalg.norm(ref) + 1e-12))
