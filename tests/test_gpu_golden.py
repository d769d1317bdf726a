"""GPU (C ABI) against the committed golden fixtures generated from the unmodified reference: the reference's
own closed-loop examples replayed with exact warm-start carry, and the seeded batches."""
import os

import numpy as np
import pytest

from conftest import assert_same
from test_golden import DT, G, batch_cases, closed_loop_replay

pytestmark = pytest.mark.gpu


def _gpu_solver(pkg, prob, dtype):
    s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")

    def solve(x0, xref, warm):
        w = {k: np.array(v, dtype=dtype, copy=True) for k, v in warm.items()}
        out = s.solve(x0, xref, warm=w)
        return out["iter"][0], out["status"][0], out["u"], out["warm"]
    return solve


@pytest.mark.parametrize("tag", list(DT))
def test_hover_closed_loop_warm_start(pkg, tag):
    rec = np.load(os.path.join(G, "hover_closed_loop_%s.npz" % tag))
    prob = pkg.problems.quadrotor(20)
    it, st, u0 = closed_loop_replay(_gpu_solver(pkg, prob, DT[tag]), prob, rec, DT[tag])
    assert_same(it, rec["iter"], "iter")
    assert_same(st, rec["status"], "status")
    assert_same(u0, rec["u0"], "u0")


@pytest.mark.parametrize("tag", list(DT))
def test_tracking_closed_loop_warm_start(pkg, tag):
    rec = np.load(os.path.join(G, "tracking_closed_loop_%s.npz" % tag))
    prob = pkg.problems.quadrotor(20)
    it, st, u0 = closed_loop_replay(_gpu_solver(pkg, prob, DT[tag]), prob, rec, DT[tag])
    assert_same(it, rec["iter"], "iter")
    assert_same(u0, rec["u0"], "u0")


def test_cartpole_closed_loop_warm_start(pkg):
    rec = np.load(os.path.join(G, "cartpole_closed_loop_f32.npz"))
    prob = pkg.problems.cartpole(max_iter=150)
    it, st, u0 = closed_loop_replay(_gpu_solver(pkg, prob, np.float32), prob, rec, np.float32)
    assert_same(it, rec["iter"], "iter")
    assert_same(u0, rec["u0"], "u0")


@pytest.mark.parametrize("tag", list(DT))
def test_seeded_batches_vs_reference_fixtures(pkg, tag):
    g = np.load(os.path.join(G, "batch_%s.npz" % tag))
    for name, prob, x0, xref in batch_cases(pkg):
        if prob.nx == 32 and tag != "f32":
            continue  # 32/8/50 is compiled for fp32 only (warp-per-instance kernel)
        out = pkg.capi.Solver(prob, dtype=DT[tag], policy="parity").solve(x0, xref)
        assert_same(out["iter"], g[name + "_iter"], name + " iter")
        assert_same(out["status"], g[name + "_status"], name + " status")
        assert_same(out["resid"], g[name + "_resid"], name + " resid")
        assert_same(out["x"][:64], g[name + "_x"], name + " x")
        assert_same(out["u"][:64], g[name + "_u"], name + " u")
        assert_same(out["x"].astype(np.float64).sum(axis=0), g[name + "_xsum"], name + " xsum")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_warm_state_written_back_exactly(pkg, oracle, dtype):
    """{d,y,g,v,z} after a solve equal the reference workspace: v/z one iteration behind on early exit, d from the
    last executed backward pass (SURVEY 8a note W), both for converged and max_iter instances; then a second,
    fully warm-started solve agrees too."""
    prob = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, 1500, mult=0.5)   # mix of converged and max_iter exits
    r1 = oracle.solve_batch(prob, x0, xref, dtype=dtype, want_state=True, nthreads=4)
    s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")
    N, nx, nu, B = prob.N, prob.nx, prob.nu, len(x0)
    warm = {k: np.zeros((B, N - 1, nu) if k in "dyz" else (B, N, nx), dtype) for k in ("d", "y", "g", "v", "z")}
    o1 = s.solve(x0, xref, warm=warm)
    assert_same(o1["iter"], r1.iter, "iter")
    assert (r1.status == 11).any() and (r1.status == 1).any()
    for k in ("d", "y", "g", "v", "z"):
        assert_same(o1["warm"][k], r1.state[k], "warm." + k)
    x0b = (x0 * 1.01).astype(np.float32)
    r2 = oracle.solve_batch(prob, x0b, xref, dtype=dtype, warm={k: r1.state[k] for k in warm}, want_state=True, nthreads=4)
    o2 = s.solve(x0b, xref, warm=o1["warm"])
    assert_same(o2["iter"], r2.iter, "iter (warm)")
    assert_same(o2["x"], r2.x, "x (warm)")
    for k in ("d", "y", "g", "v", "z"):
        assert_same(o2["warm"][k], r2.state[k], "warm2." + k)


def test_edge_cases(pkg, oracle):
    prob = pkg.problems.quadrotor(20)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    # empty batch
    out = s.solve(np.zeros((0, 12), np.float32), np.zeros((10, 12), np.float32))
    assert out["iter"].shape == (0,)
    # ragged sizes around warp / block / grid boundaries
    for B in (1, 31, 33, 129, 148 * 128 + 1):
        x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
        ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=4)
        out = s.solve(x0, xref)
        assert_same(out["iter"], ref.iter, "iter B=%d" % B)
        assert_same(out["x"], ref.x, "x B=%d" % B)
    # x0 outside the state box never converges; inverted bounds clamp to max (C.6)
    x0 = np.zeros((4, 12), np.float32); x0[:, 0] = 7.0
    ref = oracle.solve_batch(prob, x0, np.zeros((10, 12), np.float32), dtype=np.float32)
    out = s.solve(x0, np.zeros((10, 12), np.float32))
    assert (out["status"] == 11).all() and (out["iter"] == 100).all()
    assert_same(out["u"], ref.u, "u infeasible")
    # settings: check_termination > 1, bounds disabled, max_iter small
    import copy
    p2 = copy.deepcopy(prob); p2.check_termination = 3; p2.en_state_bound = 0; p2.max_iter = 17
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, 500, mult=0.25)
    ref = oracle.solve_batch(p2, x0, xref, dtype=np.float32, nthreads=4)
    out = pkg.capi.Solver(p2, dtype=np.float32, policy="parity").solve(x0, xref)
    assert_same(out["iter"], ref.iter, "iter ct=3")
    assert_same(out["resid"], ref.resid, "resid ct=3")
    assert_same(out["x"], ref.x, "x ct=3")
    with pytest.raises(pkg.capi.TmpcError):
        p3 = copy.deepcopy(prob); p3.check_termination = 0
        pkg.capi.Solver(p3)
