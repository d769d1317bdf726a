"""GPU parity of the run-time-shape kernel (csrc/tmpc_kernel_rt.cuh + tmpc_orders_rt.hpp) through the C ABI: any
nx, nu <= 64 and any horizon, bit-exact (PARITY policy) against
  * fixtures generated from the UNMODIFIED reference compiled for each shape (tests/golden/shapes_*.npz), cold and warm;
  * the CPU oracle on larger seeded batches (ragged sizes, per-instance Xref, settings variants);
  * the specialised kernels and the reference fixtures of the three BASELINE shapes (TMPC_KERNEL=rt forces this kernel).
FAST policy (FMA chains): x/u within 1e-4 relative on instances whose iteration count agrees (tolerance of north_star)."""
import os

import numpy as np
import pytest

from conftest import assert_same
from test_golden_shapes import DT, fixture_problem, load_shapes

pytestmark = pytest.mark.gpu


def _cmp_exact(out, ref, what=""):
    for name in ("iter", "status", "x", "u", "resid"):
        assert_same(out[name], getattr(ref, name), what + name)


@pytest.mark.parametrize("tag", list(DT))
def test_generic_shapes_match_reference_fixture(pkg, tag):
    z, shapes = load_shapes(tag)
    for key in shapes:
        prob = fixture_problem(pkg, z, key)
        g = lambda k: z[key + "/" + k]
        s = pkg.capi.Solver(prob, dtype=DT[tag], policy="parity")
        nx, nu, N = prob.nx, prob.nu, prob.N
        B = g("x0").shape[0]
        warm = {k: np.zeros((B, N - 1, nu) if k in "dyz" else (B, N, nx), DT[tag]) for k in ("d", "y", "g", "v", "z")}
        o1 = s.solve(g("x0"), g("xref"), warm=warm)
        for name in ("iter", "status", "resid", "x", "u"):
            assert_same(o1[name], g(name), "%s cold %s" % (key, name))
        assert s.stats()["parity_pinned"] == 1
        o2 = s.solve((g("x0") * np.float32(1.01)).astype(np.float32), g("xref"), warm=o1["warm"])
        for name in ("iter", "status", "resid", "x", "u"):
            assert_same(o2[name], g("w_" + name), "%s warm %s" % (key, name))
        for k in ("d", "y", "g", "v", "z"):
            assert_same(o2["warm"][k], g("w_state_" + k), "%s warm state %s" % (key, k))
        s.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,B", [((6, 3, 20), 4099), ((9, 4, 7), 1000), ((13, 7, 6), 777), ((20, 8, 12), 513), ((5, 1, 6), 2049),
                                     ((24, 6, 10), 300), ((2, 2, 3), 129), ((1, 1, 2), 33)])
def test_generic_shape_vs_oracle_per_instance_xref(pkg, oracle, shape, B, dtype):
    nx, nu, N = shape
    prob = pkg.problems.random_system(nx, nu, N, seed=7 + nx)
    rng = np.random.default_rng(nx * 100 + nu)
    x0 = rng.uniform(-2, 2, (B, nx)).astype(np.float32)
    x0[::2] *= np.float32(0.1)
    xref = rng.uniform(-0.3, 0.3, (B, N, nx)).astype(np.float32)
    ref = oracle.solve_batch(prob, x0, xref, dtype=dtype, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")
    out = s.solve(x0, xref)
    _cmp_exact(out, ref, "%s " % (shape,))
    st = s.stats()
    assert st["instances"] == B and st["iterations"] == int(ref.iter.sum()) and st["solved"] == int((ref.status == 1).sum())
    assert len(set(ref.iter.tolist())) > 1


def test_generic_shape_settings_variants(pkg, oracle):
    base = dict(nx=10, nu=3, N=9)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(-1.5, 1.5, (600, 10)).astype(np.float32)
    xref = np.zeros((9, 10), np.float32)
    for kw in (dict(check_termination=3), dict(max_iter=7), dict(en_state_bound=0), dict(en_input_bound=0, en_state_bound=0),
               dict(abs_pri_tol=1e-2, abs_dua_tol=1e-1)):
        prob = pkg.problems.random_system(base["nx"], base["nu"], base["N"], seed=11)
        for k, v in kw.items():
            setattr(prob, k, v)
        ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=4)
        out = pkg.capi.Solver(prob, dtype=np.float32, policy="parity").solve(x0, xref)
        _cmp_exact(out, ref, "%r " % (kw,))


@pytest.mark.parametrize("which", ["q", "c", "l"])
def test_rt_kernel_reproduces_baseline_shapes(pkg, oracle, which, monkeypatch):
    """TMPC_KERNEL=rt runs the run-time-shape kernel on the three BASELINE shapes: it must agree bit for bit with the
    oracle (hence with the specialised kernels and the reference fixtures those are pinned to), cold and warm."""
    if which == "q":
        prob = pkg.problems.quadrotor(20)
        x0, xref = pkg.workloads.quadrotor_hover_batch(0, 3000, mult=0.5)
    elif which == "c":
        prob = pkg.problems.cartpole()
        x0, xref = pkg.workloads.cartpole_batch(0, 5000)
    else:
        prob = pkg.problems.random_system()
        rng = np.random.default_rng(5)
        x0, xref = rng.uniform(-1, 1, (96, 32)).astype(np.float32), np.zeros((50, 32), np.float32)
    B = x0.shape[0]
    spec = pkg.capi.Solver(prob, dtype=np.float32, policy="parity").solve(x0, xref)
    monkeypatch.setenv("TMPC_KERNEL", "rt")
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    warm = {k: np.zeros((B, prob.N - 1, prob.nu) if k in "dyz" else (B, prob.N, prob.nx), np.float32) for k in ("d", "y", "g", "v", "z")}
    out = s.solve(x0, xref, warm=warm)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, want_state=True, nthreads=8)
    _cmp_exact(out, ref, which + " rt ")
    for name in ("iter", "status", "x", "u", "resid"):
        assert_same(out[name], spec[name], which + " rt vs specialised " + name)
    for k in ("d", "y", "g", "v", "z"):
        assert_same(out["warm"][k], ref.state[k], which + " warm state " + k)


def test_generic_shape_fast_policy_tolerance(pkg, oracle):
    prob = pkg.problems.random_system(10, 4, 12, seed=21)
    rng = np.random.default_rng(9)
    x0 = (rng.uniform(-1, 1, (4000, 10)) * 0.3).astype(np.float32)
    xref = np.zeros((12, 10), np.float32)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)
    out = pkg.capi.Solver(prob, dtype=np.float32, policy="fast").solve(x0, xref)
    same = out["iter"] == ref.iter
    assert same.mean() > 0.85, "iteration counts agree on %.3f" % same.mean()
    scale = np.abs(ref.x[same]).max(axis=(1, 2), keepdims=True) + 1e-6
    assert (np.abs(out["x"][same] - ref.x[same]) / scale).max() < 1e-4       # north_star: 1e-4 relative in fp32
    scale_u = np.abs(ref.u[same]).max(axis=(1, 2), keepdims=True) + 1e-6
    assert (np.abs(out["u"][same] - ref.u[same]) / scale_u).max() < 1e-4


def test_generic_shape_device_memory_path(pkg, oracle):
    import torch
    prob = pkg.problems.random_system(7, 2, 11, seed=4)
    B = 3001
    rng = np.random.default_rng(1)
    x0 = (rng.uniform(-1, 1, (B, 7)) * 0.5).astype(np.float32)
    xref = np.zeros((11, 7), np.float32)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    dev = torch.device("cuda:0")
    tx0, txr = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
    x = torch.empty((B, 11, 7), device=dev); u = torch.empty((B, 10, 2), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); stt = torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), device=dev)
    s.solve_raw(B, tx0, txr, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, stt, rs, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_same(it.cpu().numpy(), ref.iter, "iter")
    assert_same(x.cpu().numpy(), ref.x, "x")
    assert_same(u.cpu().numpy(), ref.u, "u")
    assert_same(rs.cpu().numpy(), ref.resid, "resid")


def test_shape_limits_are_reported(pkg):
    prob = pkg.problems.random_system(8, 2, 5, seed=2)
    prob.nx = 65
    with pytest.raises(pkg.capi.TmpcError):
        pkg.capi.Solver(prob, dtype=np.float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(6, 3, 7), (9, 4, 7), (13, 7, 6), (16, 8, 5), (5, 1, 6), (24, 6, 4)])
def test_generic_shape_step_functions_vs_oracle(pkg, oracle, shape, dtype):
    """The six step functions (admm.hpp:13-18) through tmpc_step on random full workspaces, bit-exact vs the oracle's."""
    nx, nu, N = shape
    prob = pkg.problems.random_system(nx, nu, N, seed=31 + nx)
    s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")
    rng = np.random.default_rng(17 * nx + nu)
    B = 5
    names = ["x", "u", "q", "r", "p", "d", "v", "vnew", "z", "znew", "g", "y"]
    for which in range(6):
        ws = {k: rng.uniform(-1, 1, (B, N, nx) if k in ("x", "q", "p", "v", "vnew", "g") else (B, N - 1, nu)).astype(dtype) for k in names}
        ws["Xref"] = rng.uniform(-1, 1, (N, nx)).astype(dtype)
        ws["resid"] = rng.uniform(0, 1, (B, 4)).astype(dtype)
        exp, exp_rc = [], []
        for b in range(B):   # oracle image: x u q r p d v vnew z znew g y Xref resid
            img = np.concatenate([ws[k][b].reshape(-1) for k in names] + [ws["Xref"].reshape(-1), ws["resid"][b]])
            rc, out = oracle.step(prob, which, img, it=1, dtype=dtype)
            exp.append(out)
            exp_rc.append(rc)
        term = s.step(which, ws, it=1)
        for b in range(B):
            got = np.concatenate([ws[k][b].reshape(-1) for k in names] + [ws["Xref"].reshape(-1), ws["resid"][b]])
            assert_same(got, exp[b], "%s step %d instance %d" % (shape, which, b))
            assert int(term[b]) == exp_rc[b]


@pytest.mark.parametrize("shape", [(6, 3, 8), (9, 2, 6), (16, 8, 6), (10, 12, 5)])
def test_generic_shape_rollout_vs_oracle(pkg, oracle, shape):
    """Closed loop on the device (window = fixed Xref, duals reset, warm solve, plant step) for shapes beyond BASELINE's."""
    nx, nu, N = shape
    prob = pkg.problems.random_system(nx, nu, N, seed=41 + nx)
    B, steps = 200, 4
    rng = np.random.default_rng(nx)
    x0 = (rng.uniform(-1, 1, (B, nx)) * 0.4).astype(np.float32)
    xref = np.zeros((N, nx), np.float32)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(x0)
    b.set_xref(xref)
    h = b.rollout(steps, reset_duals=True)
    warm = {k: np.zeros((B, N - 1, nu) if k in "dyz" else (B, N, nx), np.float32) for k in ("d", "y", "g", "v", "z")}
    x = x0
    for k in range(steps):
        warm["y"][:] = 0
        warm["g"][:] = 0
        r = oracle.solve_batch(prob, x, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=8)
        warm = {kk: r.state[kk] for kk in warm}
        assert_same(h["iter"][k], r.iter, "iter step %d" % k)
        assert_same(h["u0"][k], r.u[:, 0], "u0 step %d" % k)
        x = oracle.plant_step(prob, x, r.u[:, 0], dtype=np.float32)
        assert_same(h["x0"][k + 1], x, "x0 step %d" % k)


SWEEP = [(1, 5, 9), (2, 16, 7), (3, 31, 4), (4, 56, 5), (5, 8, 3), (7, 32, 5), (8, 8, 2), (9, 24, 2), (11, 56, 3), (15, 1, 11),
         (18, 20, 3), (19, 13, 7), (22, 16, 11), (28, 32, 2), (32, 31, 6), (33, 12, 9), (45, 64, 4), (55, 2, 5), (56, 7, 2),
         (63, 12, 7), (64, 15, 2), (64, 64, 4), (1, 64, 3), (1, 1, 4), (57, 57, 3)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_random_shape_sweep_vs_oracle(pkg, oracle, dtype):
    """Shapes drawn across the whole supported range (every one of them also pinned oracle-vs-reference by
    oracle/pin_shapes.py in the build container): cold solve + warm-started re-solve with state write-back."""
    for shape in SWEEP:
        nx, nu, N = shape
        prob = pkg.problems.random_system(nx, nu, N, seed=3 + nx + 64 * nu)
        rng = np.random.default_rng(nx * 64 + nu)
        B = 70
        x0 = rng.uniform(-2, 2, (B, nx)).astype(np.float32)
        x0[::2] *= np.float32(0.05)
        xref = rng.uniform(-0.2, 0.2, (N, nx)).astype(np.float32)
        s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")
        warm = {k: np.zeros((B, N - 1, nu) if k in "dyz" else (B, N, nx), dtype) for k in ("d", "y", "g", "v", "z")}
        o1 = s.solve(x0, xref, warm=warm)
        r1 = oracle.solve_batch(prob, x0, xref, dtype=dtype, want_state=True, nthreads=4)
        _cmp_exact(o1, r1, "%s cold " % (shape,))
        x1 = (x0 * np.float32(0.97)).astype(np.float32)
        o2 = s.solve(x1, xref, warm=o1["warm"])
        r2 = oracle.solve_batch(prob, x1, xref, dtype=dtype, warm={k: r1.state[k] for k in warm}, want_state=True, nthreads=4)
        _cmp_exact(o2, r2, "%s warm " % (shape,))
        for k in warm:
            assert_same(o2["warm"][k], r2.state[k], "%s warm state %s" % (shape, k))
        s.close()
