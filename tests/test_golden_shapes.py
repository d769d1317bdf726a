"""Shapes beyond the three BASELINE ones: the plain-C oracle against fixtures generated from the UNMODIFIED reference
compiled for each shape (oracle/make_golden_shapes.py -> tests/golden/shapes_{f32,f64}.npz).  This pins the oracle's
shape-generic evaluation-order dispatch (select_orders, incl. the address-peeled rows) on machines where
/root/reference is absent; test_oracle_vs_ref.py::test_generic_shapes does it live where it is present."""
import os

import numpy as np
import pytest

from conftest import assert_same

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"f32": np.float32, "f64": np.float64}


def load_shapes(tag):
    z = np.load(os.path.join(GOLD, "shapes_%s.npz" % tag))
    shapes = sorted({k.split("/")[0] for k in z.files}, key=lambda s: tuple(int(t) for t in s.split("_")))
    return z, shapes


def fixture_problem(pkg, z, key):
    nx, nu, N = (int(t) for t in key.split("_"))
    g = lambda k: z[key + "/" + k]
    return pkg.problems.Problem(nx=nx, nu=nu, N=N, rho=float(g("rho")), Adyn=g("Adyn"), Bdyn=g("Bdyn"), Q=g("Q"), Kinf=g("Kinf"),
                                Pinf=g("Pinf"), Quu_inv=g("Quu_inv"), AmBKt=g("AmBKt"), x_min=g("x_min"), x_max=g("x_max"),
                                u_min=g("u_min"), u_max=g("u_max"), name="fixture_" + key)


@pytest.mark.parametrize("tag", list(DT))
def test_oracle_matches_reference_fixture_for_generic_shapes(pkg, oracle, tag):
    z, shapes = load_shapes(tag)
    assert len(shapes) >= 20
    for key in shapes:
        prob = fixture_problem(pkg, z, key)
        g = lambda k: z[key + "/" + k]
        o1 = oracle.solve_batch(prob, g("x0"), g("xref"), dtype=DT[tag], want_state=True, nthreads=2)
        for name in ("iter", "status", "resid", "x", "u"):
            assert_same(getattr(o1, name), g(name), "%s cold %s" % (key, name))
        warm = {k: o1.state[k] for k in ("d", "y", "g", "v", "z")}
        o2 = oracle.solve_batch(prob, (g("x0") * np.float32(1.01)).astype(np.float32), g("xref"), dtype=DT[tag], warm=warm,
                                want_state=True, nthreads=2)
        for name in ("iter", "status", "resid", "x", "u"):
            assert_same(getattr(o2, name), g("w_" + name), "%s warm %s" % (key, name))
        for k in ("d", "y", "g", "v", "z"):
            assert_same(o2.state[k], g("w_state_" + k), "%s warm state %s" % (key, k))
