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


def feat_err(got, ref):
    """Largest relative feature error over every band that carries signal: |got - ref| / |ref| wherever the oracle's
    value exceeds 1e-12 x the group's largest feature.  No tolerance floor - a noise-floor band 1e-6 below the occupied
    channel is held to the same relative bound as the channel itself."""
    import numpy as np
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    live = np.abs(ref) > 1e-12 * np.abs(ref).max(axis=-1, keepdims=True)
    if not live.any():
        return 0.0
    return float((np.abs(got - ref)[live] / np.abs(ref)[live]).max())


def feat_close(got, ref, rtol):
    """Feature parity (BASELINE: channel energies <= 1e-4 relative): relative error <= rtol on every band whose oracle
    value exceeds 1e-12 x the group's largest feature (feat_err); below that - bands that hold nothing but
    single-precision rounding noise, e.g. the empty bands of a pure tone, ~1e-16 of the occupied one, which have no
    significant digits in ANY fp32 FFT - only an absolute bound of 1e-7 x the largest feature applies."""
    import numpy as np
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    top = np.abs(ref).max(axis=-1, keepdims=True)
    live = np.abs(ref) > 1e-12 * top
    ok_live = np.abs(got - ref) <= rtol * np.abs(ref)
    ok_dead = np.abs(got - ref) <= 1e-7 * top
    return bool(np.all(np.where(live, ok_live, ok_dead)))
