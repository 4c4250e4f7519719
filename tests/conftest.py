import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def O():
    """CPU oracle (test infrastructure, oracle/)."""
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def built_lib():
    """libo3d_b200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from osinco3d_b200 import build as b
    b.build()
    import osinco3d_b200
    osinco3d_b200.lib()
    return osinco3d_b200


@pytest.fixture(scope="session")
def gpu(built_lib):
    """The product library on a real device.  GPU tests FAIL (not skip) without a device:
    there is no CPU fallback to hide behind."""
    n = built_lib.device_count()
    assert n > 0, "no CUDA device: -m gpu tests must run on the GPU box"
    return built_lib


def rand_field(shape, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(scale * rng.standard_normal(shape))


def smooth_field(shape, seed):
    """smooth, non-symmetric field (so parity mistakes show) with O(1) derivatives"""
    nx, ny, nz = shape
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, nx)[:, None, None]
    y = np.linspace(0.0, 1.0, ny)[None, :, None]
    z = np.linspace(0.0, 1.0, nz)[None, None, :]
    a = rng.uniform(0.5, 3.0, 6)
    f = np.sin(a[0] * x + 0.3) * np.cos(a[1] * y - 0.2) * np.sin(a[2] * z + 0.7) + \
        0.5 * np.cos(a[3] * x * y) + 0.25 * np.sin(a[4] * y * z + a[5] * x)
    return np.asfortranarray(f)


def bits_equal(a, b):
    """bitwise equality up to the sign of zero"""
    return np.array_equal(a, b)


def rel_max(a, b):
    den = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / den)
