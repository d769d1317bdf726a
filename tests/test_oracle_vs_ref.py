"""Pins the plain-C oracle (oracle/tinympc_oracle.c) to the UNMODIFIED reference compiled from
/root/reference (oracle/_ref, built by oracle/Makefile): bit-for-bit on full solves (whole workspace) and on
each of the six step functions (admm.hpp:13-18), for every shipped shape, float and double.
Skipped where oracle/_ref is absent; the committed fixtures (test_golden.py) cover that case."""
import numpy as np
import pytest

from conftest import assert_same
from oracle.pyoracle import RefLib, ws_size

SHAPES = ["q", "c", "l"]
DT = {"f32": np.float32, "f64": np.float64}


def _prob(pkg, shape):
    return {"q": pkg.problems.quadrotor, "c": pkg.problems.cartpole, "l": pkg.problems.random_system}[shape]()


def _inputs(pkg, shape, B):
    if shape == "q":
        return pkg.workloads.quadrotor_hover_batch(0, B, mult=0.5)
    if shape == "c":
        return pkg.workloads.cartpole_batch(0, B)
    rng = np.random.default_rng(5)
    return rng.uniform(-1, 1, (B, 32)).astype(np.float32), np.zeros((50, 32), np.float32)


@pytest.mark.parametrize("tag", list(DT))
@pytest.mark.parametrize("shape", SHAPES)
def test_full_solve_whole_workspace(pkg, oracle, shape, tag):
    cfg = "%s_%s" % (shape, tag)
    if not RefLib.available(cfg):
        pytest.skip("oracle/_ref/libref_%s.so not built" % cfg)
    prob = _prob(pkg, shape)
    x0, xref = _inputs(pkg, shape, {"q": 3000, "c": 3000, "l": 120}[shape])
    r = RefLib(cfg).solve_batch(prob, x0, xref, want_state=True, nthreads=4)
    o = oracle.solve_batch(prob, x0, xref, dtype=DT[tag], want_state=True, nthreads=4)
    for name in ("iter", "status", "x", "u", "resid"):
        assert_same(getattr(o, name), getattr(r, name), name)
    for k in r.state:
        assert_same(o.state[k], r.state[k], "workspace." + k)


@pytest.mark.parametrize("tag", list(DT))
@pytest.mark.parametrize("shape", SHAPES)
def test_warm_start_chain(pkg, oracle, shape, tag):
    """Two chained solves carrying {d,y,g,v,z}: the second one starts from the exact workspace of the first."""
    cfg = "%s_%s" % (shape, tag)
    if not RefLib.available(cfg):
        pytest.skip("oracle/_ref/libref_%s.so not built" % cfg)
    prob = _prob(pkg, shape)
    x0, xref = _inputs(pkg, shape, {"q": 400, "c": 400, "l": 24}[shape])
    ref = RefLib(cfg)
    r1 = ref.solve_batch(prob, x0, xref, want_state=True)
    o1 = oracle.solve_batch(prob, x0, xref, dtype=DT[tag], want_state=True)
    x0b = (x0 * 1.01).astype(np.float32)
    wr = {k: r1.state[k] for k in ("d", "y", "g", "v", "z")}
    wo = {k: o1.state[k] for k in ("d", "y", "g", "v", "z")}
    r2 = ref.solve_batch(prob, x0b, xref, warm=wr, want_state=True)
    o2 = oracle.solve_batch(prob, x0b, xref, dtype=DT[tag], warm=wo, want_state=True)
    for name in ("iter", "status", "x", "u"):
        assert_same(getattr(o2, name), getattr(r2, name), name)
    for k in ("d", "y", "g", "v", "z"):
        assert_same(o2.state[k], r2.state[k], "warm." + k)


@pytest.mark.parametrize("tag", list(DT))
@pytest.mark.parametrize("shape", SHAPES)
def test_step_functions(pkg, oracle, shape, tag):
    cfg = "%s_%s" % (shape, tag)
    if not RefLib.available(cfg):
        pytest.skip("oracle/_ref/libref_%s.so not built" % cfg)
    prob = _prob(pkg, shape)
    ref = RefLib(cfg)
    rng = np.random.default_rng(3)
    n = ws_size(prob.nx, prob.nu, prob.N)
    for trial in range(20 if shape != "l" else 4):
        ws = rng.uniform(-1, 1, n).astype(DT[tag])
        for which in range(6):
            rc_r, out_r = ref.step(prob, which, ws, it=1)
            rc_o, out_o = oracle.step(prob, which, ws, it=1, dtype=DT[tag])
            assert rc_r == rc_o
            assert_same(out_o, out_r, "step %d" % which)


@pytest.mark.parametrize("tag", list(DT))
@pytest.mark.parametrize("shape", SHAPES)
def test_plant_step(pkg, oracle, shape, tag):
    """The examples' plant step `x1 = Adyn * x0 + Bdyn * u.col(0)` (quadrotor_hovering.cpp:108) on the reference's
    Eigen types against the oracle's restatement (= one forward_pass stage)."""
    cfg = "%s_%s" % (shape, tag)
    if not RefLib.available(cfg):
        pytest.skip("oracle/_ref/libref_%s.so not built" % cfg)
    prob = _prob(pkg, shape)
    ref = RefLib(cfg)
    rng = np.random.default_rng(11)
    B = 64
    x0 = rng.uniform(-2, 2, (B, prob.nx)).astype(DT[tag])
    u = rng.uniform(-1, 1, (B, prob.N - 1, prob.nu)).astype(DT[tag])
    exp = np.stack([ref.plant_step(prob, x0[b], u[b]) for b in range(B)])
    got = oracle.plant_step(prob, x0, u[:, 0], dtype=DT[tag])
    assert_same(got, exp, "plant step")


@pytest.mark.parametrize("shape", [(5, 2, 6), (9, 4, 7), (13, 7, 6), (24, 6, 10), (12, 9, 5), (5, 1, 6)])
def test_generic_shapes(pkg, oracle, shape):
    """Shapes beyond the three BASELINE ones: the reference is compiled for the shape on the spot (oracle/Makefile
    refshape) and the oracle's shape-generic evaluation-order dispatch must match it bit for bit -- full solves, every
    step function and the examples' plant step, float and double (oracle/pin_shapes.py runs the same check on 60+ shapes)."""
    import os
    if not os.path.isdir("/root/reference/src/tinympc"):
        pytest.skip("/root/reference not present (tests/test_golden_shapes.py covers the committed fixtures)")
    from oracle import pin_shapes
    for sc in ("f32", "f64"):
        bad, mean_it, solved = pin_shapes.check(pkg, oracle, shape, sc)
        assert not bad, "%s %s: %s" % (shape, sc, bad)
