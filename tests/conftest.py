import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=""):
    """The parity bar of BASELINE.json north_star: 1e-5 relative in fp32, with an absolute
    floor for outputs near zero (hard parts)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    if a.size == 0:
        return
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), a.shape)
        raise AssertionError("%s: %d/%d out of tolerance; worst at %s: got %r want %r" %
                             (what, int(bad.sum()), a.size, i, a[i], b[i]))
