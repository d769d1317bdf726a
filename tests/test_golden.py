"""The CPU oracle against the committed golden fixtures (tests/golden/*.npz, generated from the unmodified
reference by oracle/make_golden.py).  These run everywhere, including the GPU box where /root/reference and
possibly oracle/_ref do not exist."""
import os

import numpy as np
import pytest

from conftest import assert_same

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"f32": np.float32, "f64": np.float64}


def closed_loop_replay(solve, prob, rec, dtype):
    """Replay a stored closed loop: feed the stored x0/Xref of every step, carry d,v,z, reset y,g
    (examples/quadrotor_hovering.cpp:95-104).  `solve(x0[1,nx], xref, warm) -> (iter, status, u[1,N-1,nu], warm')`."""
    nx, nu, N = prob.nx, prob.nu, prob.N
    warm = {k: np.zeros((1, N - 1, nu) if k in "dyz" else (1, N, nx), dtype) for k in ("d", "y", "g", "v", "z")}
    iters, stats, u0s = [], [], []
    for k in range(len(rec["iter"])):
        warm["y"][:] = 0
        warm["g"][:] = 0
        it, st, u, warm = solve(rec["x0"][k][None, :], rec["xref"][k], warm)
        iters.append(int(it)); stats.append(int(st)); u0s.append(u[0, 0].copy())
    return np.array(iters), np.array(stats), np.array(u0s)


def _oracle_solver(oracle, prob, dtype):
    def solve(x0, xref, warm):
        r = oracle.solve_batch(prob, x0, xref, dtype=dtype, warm=warm, want_state=True)
        return r.iter[0], r.status[0], r.u, {k: r.state[k] for k in ("d", "y", "g", "v", "z")}
    return solve


@pytest.mark.parametrize("tag", list(DT))
def test_hover_closed_loop(pkg, oracle, tag):
    """G1-G3 of SURVEY 4.2: 70 MPC steps of examples/quadrotor_hovering.cpp."""
    rec = np.load(os.path.join(G, "hover_closed_loop_%s.npz" % tag))
    prob = pkg.problems.quadrotor(20)
    it, st, u0 = closed_loop_replay(_oracle_solver(oracle, prob, DT[tag]), prob, rec, DT[tag])
    assert_same(it, rec["iter"], "iter")
    assert_same(st, rec["status"], "status")
    assert_same(u0, rec["u0"], "u0")
    if tag == "f64":  # the sequence the survey captured from the reference binary
        assert list(it[:16]) == [100] * 8 + [31, 19, 21, 38, 38, 31, 20, 12] and list(it[-4:]) == [2, 2, 1, 2]
        assert list(st[:9]) == [11] * 8 + [1]


@pytest.mark.parametrize("tag", list(DT))
def test_tracking_closed_loop(pkg, oracle, tag):
    """G4: 290 MPC steps of examples/quadrotor_tracking.cpp."""
    rec = np.load(os.path.join(G, "tracking_closed_loop_%s.npz" % tag))
    prob = pkg.problems.quadrotor(20)
    it, st, u0 = closed_loop_replay(_oracle_solver(oracle, prob, DT[tag]), prob, rec, DT[tag])
    assert_same(it, rec["iter"], "iter")
    assert_same(u0, rec["u0"], "u0")
    if tag == "f64":
        assert it[0] == 15 and (it[1:] == 10).all() and (st == 1).all()


def test_cartpole_closed_loop(pkg, oracle):
    """G5: generated-code cartpole loop (examples/codegen_cartpole.cpp:75-122), float, max_iter 150."""
    rec = np.load(os.path.join(G, "cartpole_closed_loop_f32.npz"))
    prob = pkg.problems.cartpole(max_iter=150)
    assert prob.extra["riccati_iters"] == 476
    np.testing.assert_allclose(prob.Kinf.reshape(-1), [-2.9121762289216022, -4.8173683953046273,
                                                        44.3538695616184029, 19.7167443997914269], rtol=1e-9)
    it, st, u0 = closed_loop_replay(_oracle_solver(oracle, prob, np.float32), prob, rec, np.float32)
    assert_same(it, rec["iter"], "iter")
    assert_same(u0, rec["u0"], "u0")
    assert np.bincount(it).tolist() == [0, 119, 176, 4, 1]


def batch_cases(pkg):
    W, P = pkg.workloads, pkg.problems
    q, c, l = P.quadrotor(20), P.cartpole(), P.random_system()
    cases = []
    for mult in (0.1, 0.25, 1.0):
        cases.append(("hover_m%s" % mult, q) + W.quadrotor_hover_batch(0, 2000, mult=mult))
    cases.append(("tracking", q) + W.quadrotor_tracking_batch(0, 1160))
    cases.append(("cartpole", c) + W.cartpole_batch(0, 4000))
    rng = np.random.default_rng(7)
    cases.append(("random32", l, rng.uniform(-1, 1, (96, 32)).astype(np.float32), np.zeros((50, 32), np.float32)))
    return cases


@pytest.mark.parametrize("tag", list(DT))
def test_seeded_batches(pkg, oracle, tag):
    """G6-style seeded batches, all three shapes."""
    g = np.load(os.path.join(G, "batch_%s.npz" % tag))
    for name, prob, x0, xref in batch_cases(pkg):
        r = oracle.solve_batch(prob, x0, xref, dtype=DT[tag], nthreads=4)
        assert_same(r.iter, g[name + "_iter"], name + " iter")
        assert_same(r.status, g[name + "_status"], name + " status")
        assert_same(r.resid, g[name + "_resid"], name + " resid")
        assert_same(r.x[:64], g[name + "_x"], name + " x")
        assert_same(r.u[:64], g[name + "_u"], name + " u")
        assert_same(r.x.astype(np.float64).sum(axis=0), g[name + "_xsum"], name + " xsum")


@pytest.mark.parametrize("tag", list(DT))
@pytest.mark.parametrize("shape", ["q", "c", "l"])
def test_step_function_vectors(pkg, oracle, shape, tag):
    g = np.load(os.path.join(G, "steps_%s_%s.npz" % (shape, tag)))
    prob = {"q": pkg.problems.quadrotor, "c": pkg.problems.cartpole, "l": pkg.problems.random_system}[shape]()
    for t, ws in enumerate(g["ws_in"]):
        for which in range(6):
            rc, out = oracle.step(prob, which, ws, it=1, dtype=DT[tag])
            assert rc == g["rc%d" % which][t]
            assert_same(out, g["out%d" % which][t], "step %d trial %d" % (which, t))


ROLLOUTS = [("hover", "f32"), ("hover", "f64"), ("tracking", "f32"), ("tracking", "f64"), ("cartpole", "f32")]


def rollout_setup(pkg, name):
    """(problem, reference table [rows, nx]) of a rollout fixture (oracle/make_golden.py gen_rollouts)."""
    if name == "cartpole":
        prob = pkg.problems.cartpole(max_iter=150)
        return prob, np.zeros((prob.N, 4))
    prob = pkg.problems.quadrotor(20)
    if name == "hover":
        return prob, np.tile(pkg.workloads.QUAD_HOVER[None, :], (prob.N, 1))
    return prob, pkg.problems.quadrotor_trajectory().T


@pytest.mark.parametrize("name,tag", ROLLOUTS)
def test_rollouts_with_reference_plant(pkg, oracle, name, tag):
    """The examples' closed loop INCLUDING the reference's own plant step (x1 = Adyn*x0 + Bdyn*u.col(0) on its Eigen
    types, ref_plant_step): the oracle's solve + plant step reproduce every state, input and iteration count."""
    rec = np.load(os.path.join(G, "rollout_%s_%s.npz" % (name, tag)))
    prob, table = rollout_setup(pkg, name)
    dt = DT[tag]
    nx, nu, N = prob.nx, prob.nu, prob.N
    B, steps = rec["iter"].shape
    x0 = rec["x0_init"].astype(dt)
    assert_same(x0, rec["x0"][:, 0], "initial state")
    warm = {k: np.zeros((B, N - 1, nu) if k in "dyz" else (B, N, nx), dt) for k in ("d", "y", "g", "v", "z")}
    for k in range(steps):
        w0 = np.minimum(rec["starts"] + k, table.shape[0] - N)
        xref = np.stack([table[w:w + N] for w in w0]).astype(dt)
        warm["y"][:] = 0
        warm["g"][:] = 0
        r = oracle.solve_batch(prob, x0, xref, dtype=dt, warm=warm, want_state=True, nthreads=4)
        warm = {kk: r.state[kk] for kk in warm}
        assert_same(r.iter, rec["iter"][:, k], "iter step %d" % k)
        assert_same(r.u[:, 0], rec["u0"][:, k], "u0 step %d" % k)
        x0 = oracle.plant_step(prob, x0, r.u[:, 0], dtype=dt)
        assert_same(x0, rec["x0"][:, k + 1], "x0 after step %d" % k)
    if name == "hover" and tag == "f64":   # G1/G2: the printed tracking error and iteration sequence of the reference binary
        assert ["%.4f" % e for e in rec["err"][0, :5]] == ["2.2472", "2.9549", "2.5478", "2.6331", "3.1375"]
        assert "%.4f" % rec["err"][0, 9] == "4.6282" and "%.4f" % rec["err"][0, 35] == "0.0217" and "%.4f" % rec["err"][0, 69] == "0.0052"
        g2 = "100 100 100 100 100 100 100 100 31 19 21 38 38 31 20 12 12 11 11 10 9 9 9 9 8 8 7 6 7 7 7 7 7 6 6 6 6 6 5 5 5 4 4 3 2 2 " \
             "3 3 3 3 3 3 3 3 2 2 2 2 2 2 2 2 2 2 2 2 2 2 1 2"
        assert list(rec["iter"][0]) == [int(t) for t in g2.split()]
