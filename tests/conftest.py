"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run in the CPU-only build container (oracle vs golden vectors, host logic, ABI exports);
`-m gpu` tests are the parity tests proper and call the CUDA path through the C ABI on a real B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree native library (built on demand here; on the GPU box the prebuilt .so travels with the snapshot)."""
    from covidcurve_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


@pytest.fixture(scope="session")
def modfit():
    from tests.golden import specs
    return specs.modfit()


@pytest.fixture(scope="session")
def simdata_binomial():
    from tests.golden import specs
    return specs.simdata(True)


@pytest.fixture(scope="session")
def simdata_logit():
    from tests.golden import specs
    return specs.simdata(False)


def relerr(a, b):
    """max relative error with exact equality required where either side is the in-band failure value"""
    import numpy as np
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    clamp = -(np.finfo(float).max / 100.0)
    bad = (a == clamp) | (b == clamp)
    assert np.array_equal(a[bad], b[bad]), "in-band failure values (-DBL_MAX/100) must agree exactly"
    ok = ~bad
    if not ok.any():
        return 0.0
    return float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)))
