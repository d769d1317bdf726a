"""GPU: the device-resident batch API (tmpc_batch_*: the reference's wrapper calls with a batch dimension, and the
examples' closed loop on the device) against the reference fixtures and the CPU oracle."""
import os

import numpy as np
import pytest

from conftest import assert_same
from test_golden import DT, G, ROLLOUTS, rollout_setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,tag", ROLLOUTS)
def test_rollout_reproduces_reference_closed_loop(pkg, name, tag):
    """tmpc_batch_rollout = quadrotor_hovering.cpp:90-114 / quadrotor_tracking.cpp:93-118 on the device: every
    plant state, first input and iteration count equals the unmodified reference's (fixtures include ITS plant step)."""
    rec = np.load(os.path.join(G, "rollout_%s_%s.npz" % (name, tag)))
    prob, table = rollout_setup(pkg, name)
    B, steps = rec["iter"].shape
    s = pkg.capi.Solver(prob, dtype=DT[tag], policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(rec["x0_init"])
    b.set_xref_table(table, rec["starts"])
    h = b.rollout(steps, reset_duals=True)
    assert_same(h["iter"].T, rec["iter"], "iter")
    assert_same(h["status"].T, rec["status"], "status")
    assert_same(np.transpose(h["u0"], (1, 0, 2)), rec["u0"], "u0")
    assert_same(np.transpose(h["x0"], (1, 0, 2)), rec["x0"], "plant states")
    assert_same(b.get("x0"), rec["x0"][:, -1], "final measurement")
    assert b.last_rollout_ms() > 0


def test_wrapper_style_calls(pkg, oracle):
    """set_x0 / set_xref / reset_dual_variables / call_tiny_solve / get_x / get_u with a batch dimension
    (tiny_wrapper.cpp:5-176): two chained solves, the second warm-started from the workspace the first left."""
    prob = pkg.problems.quadrotor(20)
    B = 700
    x0, _ = pkg.workloads.quadrotor_tracking_batch(0, B)
    _, xref = pkg.workloads.quadrotor_tracking_batch(0, B)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(x0)
    b.set_xref(xref)
    b.solve()
    r1 = oracle.solve_batch(prob, x0, xref, dtype=np.float32, want_state=True, nthreads=8)
    for k in ("iter", "status", "x", "u", "resid"):
        assert_same(b.get(k), getattr(r1, k), k)
    for k in ("d", "y", "g", "v", "z"):
        assert_same(b.get(k), r1.state[k], "workspace." + k)
    assert s.stats()["iterations"] == int(r1.iter.sum())
    # second call: new measurement, duals reset, d/v/z carried (what the examples do every step)
    x1 = (x0 * 1.02).astype(np.float32)
    b.set_x0(x1)
    b.reset_dual_variables()
    b.solve()
    warm = {k: r1.state[k].copy() for k in ("d", "y", "g", "v", "z")}
    warm["y"][:] = 0
    warm["g"][:] = 0
    r2 = oracle.solve_batch(prob, x1, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
    for k in ("iter", "x", "u"):
        assert_same(b.get(k), getattr(r2, k), k + " (2nd)")
    # shared Xref + cold reset
    b.reset()
    b.set_xref(xref[0])
    b.set_x0(x0)
    b.solve()
    r3 = oracle.solve_batch(prob, x0, xref[0], dtype=np.float32, nthreads=8)
    assert_same(b.get("iter"), r3.iter, "iter (shared xref, cold)")
    assert_same(b.get("u"), r3.u, "u (shared xref, cold)")


def test_rollout_large_batch_against_oracle(pkg, oracle):
    """A 3000-instance hover rollout without duals reset (full warm start carried step to step) vs the oracle."""
    prob = pkg.problems.quadrotor(20)
    B, steps = 3000, 6
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.5)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(x0)
    b.set_xref(xref)
    h = b.rollout(steps, reset_duals=False)
    warm = {k: np.zeros((B, prob.N - 1, prob.nu) if k in "dyz" else (B, prob.N, prob.nx), np.float32) for k in ("d", "y", "g", "v", "z")}
    x = x0
    for k in range(steps):
        r = oracle.solve_batch(prob, x, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
        warm = {kk: r.state[kk] for kk in warm}
        assert_same(h["iter"][k], r.iter, "iter step %d" % k)
        assert_same(h["u0"][k], r.u[:, 0], "u0 step %d" % k)
        x = oracle.plant_step(prob, x, r.u[:, 0], dtype=np.float32)
        assert_same(h["x0"][k + 1], x, "x0 step %d" % k)


def test_rollout_large_shape(pkg, oracle):
    prob = pkg.problems.random_system()
    B, steps = 24, 3
    x0, xref = pkg.workloads.random_system_batch(0, B, amp=0.5)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(x0)
    b.set_xref(xref)
    h = b.rollout(steps, reset_duals=True)
    warm = {k: np.zeros((B, prob.N - 1, prob.nu) if k in "dyz" else (B, prob.N, prob.nx), np.float32) for k in ("d", "y", "g", "v", "z")}
    x = x0
    for k in range(steps):
        warm["y"][:] = 0
        warm["g"][:] = 0
        r = oracle.solve_batch(prob, x, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
        warm = {kk: r.state[kk] for kk in warm}
        assert_same(h["iter"][k], r.iter, "iter step %d" % k)
        x = oracle.plant_step(prob, x, r.u[:, 0], dtype=np.float32)
        assert_same(h["x0"][k + 1], x, "x0 step %d" % k)
