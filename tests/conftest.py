import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def crn():
    """The product package (ctypes over libcrnsense.so); built on demand here, prebuilt on the GPU box."""
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "cognitive-radio-network_b200", "libcrnsense.so")
    if not os.path.exists(lib):
        g.build()
    import crn_b200
    return crn_b200


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.port()
    return o


def feat_close(got, ref, rtol):
    """Feature parity: |got - ref| <= rtol*|ref|, plus a floor of 1e-7 x the group's largest feature for
    bands that hold nothing but single-precision rounding noise (e.g. a pure tone: the empty bands are
    ~1e-16 of the occupied one and have no significant digits in ANY fp32 FFT)."""
    import numpy as np
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    floor = 1e-7 * np.abs(ref).max(axis=-1, keepdims=True)
    return bool(np.all(np.abs(got - ref) <= rtol * np.abs(ref) + floor))
