"""Fused closed loop of the fp32 12/4/10 kernel (ROLL instances, tmpc_kernel_f32.cuh): tmpc_batch_rollout with reset duals and a
fixed reference runs ALL the MPC steps of an instance on one lane of ONE persistent launch -- plant step, dual reset and the warm
d / v / z hand-over (one iteration behind after an early exit, admm.cpp:135-144) never leave the chip.  Everything the loop
produces must equal the oracle's closed loop (examples/quadrotor_hovering.cpp:90-114 with the reference's own plant step) bit for
bit, and the one-launch-per-step path (TMPC_ROLL=0) on the whole batch."""
import dataclasses
import os

import numpy as np
import pytest

from conftest import assert_same

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run(pkg, prob, x0, xref, steps, **env):
    with _env(**env):
        s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
        b = pkg.capi.Batch(s, x0.shape[0])
        b.set_x0(x0)
        b.set_xref(xref)
        h = b.rollout(steps, reset_duals=True)
        out = {k: b.get(k) for k in ("x", "u", "iter", "status", "resid", "x0", "d", "y", "g", "v", "z")}
        st = s.stats()
        b.close()
        s.close()
    return h, out, st


def _oracle_loop(oracle, prob, x0, xref, steps):
    hist = {"iter": [], "status": [], "u0": [], "x0": [x0.copy()]}
    xc, warm, r = x0.copy(), None, None
    for _ in range(steps):
        r = oracle.solve_batch(prob, xc, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
        hist["iter"].append(r.iter); hist["status"].append(r.status); hist["u0"].append(r.u[:, 0, :])
        xc = oracle.plant_step(prob, xc, r.u[:, 0, :], dtype=np.float32)
        hist["x0"].append(xc)
        warm = {q: r.state[q].copy() for q in ("d", "y", "g", "v", "z")}
        warm["y"][:] = 0
        warm["g"][:] = 0
    return {k: np.stack(v) for k, v in hist.items()}, r


def _check(h, out, ho, r, n, what):
    for k in ("iter", "status", "u0", "x0"):
        assert_same(h[k][:, :n], ho[k], "%s: history %s" % (what, k))
    for k in ("x", "u", "iter", "status", "resid"):
        assert_same(out[k][:n], getattr(r, k), "%s: last step %s" % (what, k))
    for k in ("d", "y", "g", "v", "z"):
        assert_same(out[k][:n], r.state[k], "%s: final workspace %s" % (what, k))
    assert_same(out["x0"][:n], ho["x0"][-1], "%s: final measurement" % what)


def _same_runs(a, b, what):
    (ha, oa, _), (hb, ob, _) = a, b
    for k in ha:
        assert_same(ha[k], hb[k], "%s: history %s" % (what, k))
    for k in oa:
        assert_same(oa[k], ob[k], "%s: %s" % (what, k))


@pytest.mark.parametrize("mode,tol", [(0, 1e-3), (1, 1e-3), (2, 1e-3), (8, 1e-3), (9, 1e-3), (0, 0.25), (1, 0.25), (8, 0.25)])
def test_fused_loop_hover_vs_oracle(pkg, oracle, mode, tol):
    """80,000 hover instances (two claims per lane and more), 5 MPC steps; the first 2,500 against the oracle.  mode 1: the mirror
    is never predicted, so every early exit past iteration 1 takes the step again on the lane; mode 2: mirror in every sweep;
    mode 8: the exact hand-over (v, z reloaded from the mirror rows) at every step instead of the lazy first iteration.
    tol 0.25: loose tolerances, so that steps end at their FIRST iteration (the state handed over is then the step's own start state)."""
    prob = dataclasses.replace(pkg.problems.quadrotor(20), abs_pri_tol=tol, abs_dua_tol=tol)
    B, steps, n = 80_000, 5, 2_500
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25, seed=3)
    x0[::3] = xref[0] + (x0[::3] - xref[0]) * np.float32(2e-3)     # a third of the batch next to the set-point: one-iteration steps
    fused = _run(pkg, prob, x0, xref, steps, TMPC_TEST_MIRROR=mode or None)
    assert fused[2]["launches"] == 2          # the fused launch + the last plant step
    ho, r = _oracle_loop(oracle, prob, x0[:n], xref, steps)
    _check(fused[0], fused[1], ho, r, n, "mode %d" % mode)
    assert (ho["iter"] > 1).any() and (tol < 0.1 or (ho["iter"] == 1).any())
    if mode == 0:
        stepwise = _run(pkg, prob, x0, xref, steps, TMPC_ROLL=0)
        assert stepwise[2]["launches"] == 2 * steps
        _same_runs(fused, stepwise, "fused vs one launch per step")


@pytest.mark.parametrize("max_iter,check", [(6, 1), (40, 3), (1, 1)])
def test_fused_loop_iteration_limits(pkg, oracle, max_iter, check):
    """max_iter exits inside the loop (their backward sweep still runs, admm.cpp:141-144, and hands d / v / z over on chip) and
    check_termination > 1."""
    prob = dataclasses.replace(pkg.problems.quadrotor(20), max_iter=max_iter, check_termination=check)
    B, steps, n = 45_000, 4, 2_000
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.4, seed=5)
    fused = _run(pkg, prob, x0, xref, steps)
    ho, r = _oracle_loop(oracle, prob, x0[:n], xref, steps)
    _check(fused[0], fused[1], ho, r, n, "max_iter %d check %d" % (max_iter, check))
    if max_iter > 1:
        assert (ho["status"] == 11).any()
    _same_runs(fused, _run(pkg, prob, x0, xref, steps, TMPC_ROLL=0), "fused vs one launch per step")


def test_fused_loop_per_instance_reference_and_dense_instance(pkg, oracle):
    """Every instance its own (fixed) reference trajectory; and the dense kernel instance (TMPC_DENSE=1: what a model that is not
    the shipped quadrotor runs on)."""
    prob = pkg.problems.quadrotor(20)
    B, steps, n = 40_000, 3, 1_500
    x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B)
    fused = _run(pkg, prob, x0, xref, steps)
    ho, r = _oracle_loop(oracle, prob, x0[:n], xref[:n], steps)
    _check(fused[0], fused[1], ho, r, n, "per-instance Xref")
    dense = _run(pkg, prob, x0, xref, steps, TMPC_DENSE=1)
    _same_runs(fused, dense, "structured vs dense instance")


def test_fused_loop_continues_a_previous_rollout(pkg, oracle):
    """Two rollouts back to back (the second starts warm from the first's workspace and is scheduled by its iteration counts)
    equal one rollout of the total length; small batches (fewer instances than lanes) included."""
    prob = pkg.problems.quadrotor(20)
    for B in (300, 160_000):
        x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25, seed=7)
        s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
        b = pkg.capi.Batch(s, B)
        b.set_x0(x0)
        b.set_xref(xref)
        h1 = b.rollout(3, reset_duals=True)
        h2 = b.rollout(2, reset_duals=True)
        if B >= 160_000:
            assert s.stats()["scheduled"] == 2     # claim order from the previous rollout's last iteration counts
        n = min(B, 1_500)
        ho, r = _oracle_loop(oracle, prob, x0[:n], xref, 5)
        for k in ("iter", "status", "u0"):
            assert_same(np.concatenate([h1[k], h2[k]])[:, :n], ho[k], "B %d history %s" % (B, k))
        assert_same(np.concatenate([h1["x0"], h2["x0"][1:]])[:, :n], ho["x0"], "B %d plant states" % B)
        for k in ("d", "y", "g", "v", "z"):
            assert_same(b.get(k)[:n], r.state[k], "B %d final workspace %s" % (B, k))
        b.close()
        s.close()


@pytest.mark.parametrize("scale,max_iter,check", [(0.2, 100, 1), (1.0, 30, 1), (0.2, 2, 1), (0.5, 100, 2), (0.2, 1, 1)])
def test_fused_loop_cartpole_registers(pkg, oracle, scale, max_iter, check):
    """4/1/10 (codegen_cartpole.cpp:75-122): the register-resident fused loop (tmpc_kernel_small.cuh admm_kernel_small_roll) keeps both
    generations of v, z, so the hand-over after an early exit is the reference's state by construction.  Steps of 1-4 iterations
    (scale 0.2), instances pinned at max_iter (scale 1.0 / max_iter 2 / 1), check_termination 2."""
    prob = dataclasses.replace(pkg.problems.cartpole(max_iter=max_iter), check_termination=check)
    B, steps, n = 120_000, 6, 3_000
    x0, xref = pkg.workloads.cartpole_batch(0, B)
    x0 = (scale * x0).astype(np.float32)
    fused = _run(pkg, prob, x0, xref, steps)
    assert fused[2]["launches"] == 2
    ho, r = _oracle_loop(oracle, prob, x0[:n], xref, steps)
    _check(fused[0], fused[1], ho, r, n, "cartpole scale %g max_iter %d check %d" % (scale, max_iter, check))
    if scale == 0.2 and max_iter == 100:
        assert (ho["iter"] == 1).any() and (ho["iter"] == 2).any()
    if max_iter <= 30:
        assert (ho["status"] == 11).any()
    _same_runs(fused, _run(pkg, prob, x0, xref, steps, TMPC_ROLL=0), "fused vs one launch per step")


def test_fused_loop_tracks_a_reference_table(pkg, oracle):
    """quadrotor_tracking.cpp:93-118 fused: every lane reads the window of its instance's current step from the table itself
    (w0 = min(start + k, rows - N): starts up to the end of the table exercise the clamp).  Against the oracle's loop with the
    windows gathered on the host, and against the one-launch-per-step path on the whole batch -- twice in a row, so the second
    rollout continues from the first one's step count."""
    prob = pkg.problems.quadrotor(20)
    table = np.ascontiguousarray(pkg.problems.quadrotor_trajectory().T).astype(np.float32)     # [301, 12]
    B, n = 50_000, 1_200
    rng = np.random.default_rng(11)
    starts = rng.integers(0, table.shape[0] - 4, B).astype(np.int32)
    x0 = (table[np.minimum(starts, table.shape[0] - prob.N)] + 0.05 * rng.standard_normal((B, 12))).astype(np.float32)

    def run(**env):
        with _env(**env):
            s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
            b = pkg.capi.Batch(s, B)
            b.set_x0(x0)
            b.set_xref_table(table, starts)
            h1 = b.rollout(4, reset_duals=True)
            l1 = s.stats()["launches"]
            h2 = b.rollout(3, reset_duals=True)
            h = {k: np.concatenate([h1[k], h2[k][1:] if k == "x0" else h2[k]]) for k in h1}
            out = {k: b.get(k) for k in ("x", "u", "iter", "status", "resid", "x0", "d", "y", "g", "v", "z")}
            b.close(); s.close()
        return h, out, l1

    fused, stepwise = run(), run(TMPC_ROLL=0)
    assert fused[2] == 2 and stepwise[2] == 3 * 4
    _same_runs(fused, stepwise, "fused vs one launch per step (table)")
    xc, warm = x0[:n].copy(), None
    for k in range(7):
        w0 = np.minimum(starts[:n].astype(np.int64) + k, table.shape[0] - prob.N)
        xref = table[w0[:, None] + np.arange(prob.N)[None, :]]
        r = oracle.solve_batch(prob, xc, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
        assert_same(fused[0]["iter"][k, :n], r.iter, "step %d iter" % k)
        assert_same(fused[0]["u0"][k, :n], r.u[:, 0, :], "step %d u0" % k)
        xc = oracle.plant_step(prob, xc, r.u[:, 0, :], dtype=np.float32)
        assert_same(fused[0]["x0"][k + 1, :n], xc, "step %d plant state" % k)
        warm = {q: r.state[q].copy() for q in ("d", "y", "g", "v", "z")}
        warm["y"][:] = 0
        warm["g"][:] = 0
    for q in ("x", "u", "iter"):
        assert_same(fused[1][q][:n], getattr(r, q), "last step " + q)
    for q in ("d", "v", "z", "y", "g"):
        assert_same(fused[1][q][:n], r.state[q], "final workspace " + q)
