import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (test infrastructure).  Built by __graft_entry__.build()."""
    from oracle.pyoracle import OracleLib
    path = os.path.join(ROOT, "oracle", "libtinympc_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return OracleLib()


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def assert_same(a, b, what=""):
    """Bit-for-bit equality of values (+0 == -0)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    bad = a != b
    if bad.any():
        idx = np.argwhere(bad)[0]
        raise AssertionError("%s: %d / %d elements differ, first at %s: %r vs %r" %
                             (what, bad.sum(), a.size, tuple(idx), a[tuple(idx)], b[tuple(idx)]))
