"""The C-ABI shared library loads without a GPU and exports every symbol include/tmpc.h declares; argument
errors are reported without touching a device; a device-needing call fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_cuda


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "tmpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tmpc_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported(pkg):
    lib = pkg.capi.load()
    names = declared_symbols()
    assert set(pkg.capi.EXPORTS) == set(names), (names, pkg.capi.EXPORTS)
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert b"sm_100a" in lib.tmpc_version()


def test_create_rejects_bad_arguments(pkg):
    lib = pkg.capi.load()
    ctx = ctypes.c_void_p()
    assert lib.tmpc_create(ctypes.byref(ctx), 0, 65, 3, 5, 0, 0) == -2     # shape outside what the kernels cover
    assert b"outside" in lib.tmpc_last_error(None)
    assert lib.tmpc_create(ctypes.byref(ctx), 0, 7, 3, 1, 0, 0) == -2      # horizon < 2
    assert lib.tmpc_create(ctypes.byref(ctx), 0, 12, 4, 10, 9, 0) == -1     # bad dtype
    assert lib.tmpc_create(None, 0, 12, 4, 10, 0, 0) == -1
    assert lib.tmpc_set_instance_bounds(None, 4, None, None, None, None, 0) == -1   # no ctx: refused before any device call


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(pkg):
    with pytest.raises(pkg.capi.TmpcError, match="no CUDA device"):
        pkg.capi.Solver(pkg.problems.quadrotor(20))
