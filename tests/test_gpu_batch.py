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


def test_argument_errors_are_reported(pkg):
    """Error behaviour of the batch / systems entry points: negative codes + a message, no crash."""
    import ctypes as C
    capi = pkg.capi
    prob = pkg.problems.quadrotor(20)
    s = capi.Solver(prob, dtype=np.float32, policy="parity")
    lib = s.lib
    h = C.c_void_p()
    assert lib.tmpc_batch_create(s._ctx, 0, C.byref(h)) == -1                       # batch < 1
    b = capi.Batch(s, 16)
    assert lib.tmpc_batch_set_x0(b._b, None, capi.TMPC_MEM_HOST) == -1              # NULL source
    assert b"NULL" in lib.tmpc_batch_last_error(b._b)
    short = np.zeros((5, 12), np.float32)
    assert lib.tmpc_batch_set_xref_table(b._b, short.ctypes.data, 5, None, capi.TMPC_MEM_HOST) == -1   # fewer rows than N
    assert lib.tmpc_batch_get(b._b, 99, short.ctypes.data, capi.TMPC_MEM_HOST) == -1                     # bad selector
    assert lib.tmpc_batch_rollout(b._b, -1, 1, None, None, None, None, capi.TMPC_MEM_HOST) == -1
    assert lib.tmpc_batch_rollout(b._b, 0, 1, None, None, None, None, capi.TMPC_MEM_HOST) == 0
    # systems: batch mismatch, host buffers rejected, wrong shape unsupported
    S = 8
    sy = capi.Systems(s, np.repeat(prob.Adyn[None], S, 0), np.repeat(prob.Bdyn[None], S, 0), np.repeat(prob.Q[None], S, 0),
                      np.repeat(prob.R[None], S, 0), np.full(S, prob.rho))
    args = capi.TmpcSolveArgs()
    args.batch = S + 1
    args.mem = capi.TMPC_MEM_DEVICE
    assert lib.tmpc_solve_systems(s._ctx, C.byref(args), sy._p) == -1
    args.batch = S
    args.mem = capi.TMPC_MEM_HOST
    assert lib.tmpc_solve_systems(s._ctx, C.byref(args), sy._p) == -1
    assert b"device buffers" in lib.tmpc_last_error(s._ctx)
    big = capi.Solver(pkg.problems.random_system(), dtype=np.float32, policy="parity")
    p = C.c_void_p()
    z = np.zeros(32 * 32 * 2, np.float32)
    assert lib.tmpc_systems_precompute(big._ctx, 2, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data, 0,
                                       capi.TMPC_MEM_HOST, C.byref(p)) == -2       # 32/8 per-instance systems not compiled
    # a singular R + B'PB is reported per instance, not as a crash
    Sz = capi.Systems(s, np.zeros((2, 12, 12)), np.zeros((2, 12, 4)), np.zeros((2, 12)), np.zeros((2, 4)), np.zeros(2))
    assert Sz.get("sweeps").tolist() == [-1, -1]


def test_large_rollout_uses_iteration_history_schedule(pkg, oracle):
    """A closed loop large enough for the iteration-history schedule (claim order = previous step's iteration counts, largest
    first) and for the in-kernel dual reset (fp32 12/4/10: y, g zero-filled on chip instead of memset + read): 160,000 hover
    instances, 4 MPC steps; the first 2,000 instances against the oracle's closed loop (its own plant step), every step."""
    prob = pkg.problems.quadrotor(20)
    B, steps, n = 160_000, 4, 2_000
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(x0)
    b.set_xref(xref)
    old = os.environ.get("TMPC_ROLL")
    os.environ["TMPC_ROLL"] = "0"      # one launch per MPC step (the fused loop has its own tests, test_gpu_roll.py)
    try:
        h = b.rollout(steps, reset_duals=True)
    finally:
        if old is None:
            del os.environ["TMPC_ROLL"]
        else:
            os.environ["TMPC_ROLL"] = old
    assert s.stats()["scheduled"] == 2
    xc, warm = x0[:n].copy(), None
    for k in range(steps):
        r = oracle.solve_batch(prob, xc, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
        assert_same(h["iter"][k, :n], r.iter, "step %d iter" % k)
        assert_same(h["status"][k, :n], r.status, "step %d status" % k)
        assert_same(h["u0"][k, :n], r.u[:, 0, :], "step %d u0" % k)
        xc = oracle.plant_step(prob, xc, r.u[:, 0, :], dtype=np.float32)
        assert_same(h["x0"][k + 1, :n], xc, "step %d plant state" % k)
        warm = {q: r.state[q].copy() for q in ("d", "y", "g", "v", "z")}
        warm["y"][:] = 0
        warm["g"][:] = 0
    # the workspace the loop leaves behind (what a following wrapper-style call would start from)
    for q in ("d", "v", "z", "y", "g"):
        assert_same(b.get(q)[:n], r.state[q], "final workspace " + q)
