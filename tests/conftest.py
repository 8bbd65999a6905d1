"""pytest configuration: `gpu` marker, repo-root imports, native / oracle libraries built on demand."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the C-ABI library and the C oracle exist (nvcc / gcc cross-compile without a GPU)."""
    from navigation_model_b200 import _build
    from oracle import c_oracle
    if _build.needs_rebuild():
        try:
            _build.build_native()
        except RuntimeError:
            if not os.path.exists(_build.LIB_PATH):
                raise
    c_oracle.build()
    yield


@pytest.fixture(scope="session")
def golden_fit():
    return dict(np.load(os.path.join(GOLDEN, "fit_golden.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_sim():
    path = os.path.join(GOLDEN, "sim_golden.npz")
    if not os.path.exists(path):
        pytest.skip("sim_golden.npz not generated")
    return dict(np.load(path, allow_pickle=False))


def has_cuda():
    try:
        from navigation_model_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda():
    if not has_cuda():
        pytest.fail("a test marked gpu ran without a usable CUDA device / native library")
    import torch
    return torch.device("cuda", 0)
