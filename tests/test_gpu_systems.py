"""GPU: per-instance SYSTEMS.  The batched device precompute against the host tiny_precompute
(bit-identical) and the reference's published cartpole cache (G5); the per-instance-model solve against the CPU
oracle run system by system on the very caches the device produced (bit-exact, PARITY policy)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, assert_same

pytestmark = pytest.mark.gpu


def _systems(pkg, S, dtype):
    """S perturbed quadrotor models: couplings of Adyn scaled, Bdyn scaled, rho and Q varied per system."""
    base = pkg.problems.quadrotor(20)
    rng = np.random.default_rng(42)
    A = np.repeat(base.Adyn[None], S, 0).copy()
    off = ~np.eye(12, dtype=bool)
    A[:, off] *= (1.0 + 0.2 * rng.uniform(-1, 1, (S, 1)))
    Bm = base.Bdyn[None] * (1.0 + 0.3 * rng.uniform(-1, 1, (S, 1, 1)))
    Q = base.Q[None] * (1.0 + 0.5 * rng.uniform(-1, 1, (S, 12)))
    R = np.repeat(base.R[None], S, 0) * (1.0 + 0.5 * rng.uniform(-1, 1, (S, 4)))
    rho = 5.0 * (1.0 + 0.4 * rng.uniform(-1, 1, S))
    return base, A.astype(dtype), Bm.astype(dtype), Q.astype(dtype), R.astype(dtype), rho.astype(dtype)


def _host_precompute(dtype, nx, nu, A, Bm, Q, R, rho):
    lib = C.CDLL(os.path.join(ROOT, "accelerated-tinympc_b200", "lib", "libtinympc_b200_%s.so" % ("f32" if dtype == np.float32 else "f64")))
    ct = C.c_float if dtype == np.float32 else C.c_double
    lib.tiny_precompute_raw.restype = C.c_int
    lib.tiny_precompute_raw.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [ct] + [C.c_void_p] * 4
    Ac = np.ascontiguousarray(A.T)   # column-major
    Bc = np.ascontiguousarray(Bm.T)
    K = np.empty((nx, nu), dtype); P = np.empty((nx, nx), dtype); Qi = np.empty((nu, nu), dtype); M = np.empty((nx, nx), dtype)
    sw = lib.tiny_precompute_raw(nx, nu, Ac.ctypes.data, Bc.ctypes.data, np.ascontiguousarray(Q).ctypes.data,
                                 np.ascontiguousarray(R).ctypes.data, ct(float(rho)), K.ctypes.data, P.ctypes.data, Qi.ctypes.data, M.ctypes.data)
    return sw, K.T, P.T, Qi.T, M.T


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_batched_precompute_equals_host(pkg, dtype):
    S = 40
    base, A, Bm, Q, R, rho = _systems(pkg, S, dtype)
    s = pkg.capi.Solver(base, dtype=dtype, policy="parity")
    sy = pkg.capi.Systems(s, A, Bm, Q, R, rho)
    sweeps = sy.get("sweeps")
    K, P, Qi, M = sy.get("Kinf"), sy.get("Pinf"), sy.get("Quu_inv"), sy.get("AmBKt")
    assert (sweeps > 10).all() and (sweeps < 1000).all()
    for i in range(S):
        sw, Kh, Ph, Qih, Mh = _host_precompute(dtype, 12, 4, A[i], Bm[i], Q[i], R[i], rho[i])
        assert sw == sweeps[i]
        assert_same(K[i], Kh, "Kinf[%d]" % i); assert_same(P[i], Ph, "Pinf[%d]" % i)
        assert_same(Qi[i], Qih, "Quu_inv[%d]" % i); assert_same(M[i], Mh, "AmBKt[%d]" % i)
    assert_same(sy.get("Adyn"), A, "Adyn"); assert_same(sy.get("rho"), rho, "rho"); assert_same(sy.get("Q"), Q, "Q")


def test_cartpole_precompute_g5(pkg):
    """The cartpole model of examples/codegen_cartpole.cpp: 476 sweeps and the cache tiny_codegen printed."""
    sc, m = pkg.problems.cartpole_model()
    prob = pkg.problems.cartpole()
    s = pkg.capi.Solver(prob, dtype=np.float64, policy="parity")
    S = 3
    sy = pkg.capi.Systems(s, np.repeat(m["Adyn"][None], S, 0), np.repeat(m["Bdyn"][None], S, 0), np.repeat(m["Q"].reshape(1, -1), S, 0),
                          np.repeat(m["R"].reshape(1, -1), S, 0), np.full(S, float(sc["rho"])), q_plus_rho=True)
    assert sy.get("sweeps").tolist() == [476] * S
    np.testing.assert_allclose(sy.get("Kinf")[0].reshape(-1), [-2.9121762289216022, -4.8173683953046273, 44.3538695616184029,
                                                               19.7167443997914269], rtol=1e-9)
    np.testing.assert_allclose(sy.get("Q")[0], m["Q"].reshape(-1) + float(sc["rho"]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_per_instance_systems_solve_bit_exact(pkg, oracle, dtype):
    import copy
    import torch
    S, per = 24, 50
    base, A, Bm, Q, R, rho = _systems(pkg, S, dtype)
    B = S * per
    idx = np.repeat(np.arange(S), per)
    s = pkg.capi.Solver(base, dtype=dtype, policy="parity")
    sy = pkg.capi.Systems(s, A[idx], Bm[idx], Q[idx], R[idx], rho[idx])
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.4)
    dev = torch.device("cuda:0")
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(dev)
    x0d, xrd = f(x0), f(xref)
    x = torch.empty((B, 10, 12), dtype=tdt, device=dev); u = torch.empty((B, 9, 4), dtype=tdt, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), dtype=tdt, device=dev)
    warm = {k: torch.zeros((B, 9, 4) if k in "dyz" else (B, 10, 12), dtype=tdt, device=dev) for k in ("d", "y", "g", "v", "z")}
    sy.solve_raw(x0d, xrd, True, x, u, it, st, rs, warm=warm)
    torch.cuda.synchronize()
    K, P, Qi, M = sy.get("Kinf"), sy.get("Pinf"), sy.get("Quu_inv"), sy.get("AmBKt")
    itn, xn, un = it.cpu().numpy(), x.cpu().numpy(), u.cpu().numpy()
    for sidx in range(S):
        p = copy.deepcopy(base)
        j = sidx * per
        p.Adyn, p.Bdyn, p.Q, p.rho = A[j // per].astype(np.float64), Bm[j // per].astype(np.float64), Q[j // per].astype(np.float64), float(rho[j // per])
        p.Kinf, p.Pinf, p.Quu_inv, p.AmBKt = (K[j].astype(np.float64), P[j].astype(np.float64), Qi[j].astype(np.float64), M[j].astype(np.float64))
        ref = oracle.solve_batch(p, x0[j:j + per], xref, dtype=dtype, want_state=True, nthreads=4)
        assert_same(itn[j:j + per], ref.iter, "iter sys %d" % sidx)
        assert_same(xn[j:j + per], ref.x, "x sys %d" % sidx)
        assert_same(un[j:j + per], ref.u, "u sys %d" % sidx)
        assert_same(warm["v"][j:j + per].cpu().numpy(), ref.state["v"], "warm v sys %d" % sidx)
    assert len(set(itn.tolist())) > 5
    stt = s.stats()
    assert stt["iterations"] == int(itn.sum()) and stt["pattern"] == 0


@pytest.mark.parametrize("S,const_bounds", [(333, True), (40_000, True), (2_000, False)])
def test_systems_kernel_variants_agree(pkg, oracle, monkeypatch, S, const_bounds):
    """fp32 12/4/10 systems solve, four kernels: two lanes per instance with row pairs streamed from tensor memory (default,
    tmpc_kernel_sysp.cuh), the same with one thread per instance (TMPC_KERNEL=sys_thread, tmpc_kernel_sys.cuh), the first
    TMEM-resident kernel (sys_rows) and coefficients re-read from the global block (sys_global) -- identical results,
    including ragged batch sizes, the refill of single lanes, the longest-first schedule (large S) and the controls-only
    output; a sample of the systems is checked against the oracle."""
    import copy
    import torch
    if not const_bounds:   # the instances of the kernels that read stage-indexed bounds (the examples' boxes are the same at every stage)
        monkeypatch.setenv("TMPC_NO_CONST_BOUNDS", "1")
    base, A, Bm, Q, R, rho = _systems(pkg, 97, np.float32)   # 97 distinct systems, repeated with different x0
    idx = np.arange(S) % 97
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, S, mult=0.6)
    dev = torch.device("cuda:0")
    outs = []
    for variant in (None, "sys_thread", "sys_rows", "sys_global"):
        if variant:
            monkeypatch.setenv("TMPC_KERNEL", variant)
        s = pkg.capi.Solver(base, dtype=np.float32, policy="parity")
        sy = pkg.capi.Systems(s, A[idx], Bm[idx], Q[idx], R[idx], rho[idx])
        x = torch.empty((S, 10, 12), device=dev); u = torch.empty((S, 9, 4), device=dev)
        it = torch.empty(S, dtype=torch.int32, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
        rs = torch.empty((S, 4), device=dev)
        u0 = torch.empty((S, 4), device=dev)
        sy.solve_raw(torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev), True, x, u, it, st, rs, u0=u0)
        torch.cuda.synchronize()
        outs.append((it.cpu().numpy(), st.cpu().numpy(), x.cpu().numpy(), u.cpu().numpy(), rs.cpu().numpy()))
        assert_same(u0.cpu().numpy(), outs[-1][3][:, 0, :], "u0 next to the full outputs (%s)" % variant)
        # controls-only output mask: no trajectories, same u0 / iter
        u0b = torch.zeros((S, 4), device=dev); itb = torch.zeros(S, dtype=torch.int32, device=dev)
        sy.solve_raw(torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev), True, None, None, itb, None, None, u0=u0b)
        torch.cuda.synchronize()
        assert_same(u0b.cpu().numpy(), outs[-1][3][:, 0, :], "u0, controls-only (%s)" % variant)
        assert_same(itb.cpu().numpy(), outs[-1][0], "iter, controls-only (%s)" % variant)
        if variant is None:
            K, P, Qi, M = sy.get("Kinf"), sy.get("Pinf"), sy.get("Quu_inv"), sy.get("AmBKt")
        if variant:
            monkeypatch.delenv("TMPC_KERNEL")
    for o in outs[1:]:
        for a, b, name in zip(outs[0], o, ("iter", "status", "x", "u", "resid")):
            assert_same(a, b, name)
    assert outs[0][0].min() < outs[0][0].max()
    for sidx in (0, 13, 96):
        sel = np.nonzero(idx == sidx)[0][:64]
        p = copy.deepcopy(base)
        j = sel[0]
        p.Adyn, p.Bdyn, p.Q, p.rho = A[sidx].astype(np.float64), Bm[sidx].astype(np.float64), Q[sidx].astype(np.float64), float(rho[sidx])
        p.Kinf, p.Pinf, p.Quu_inv, p.AmBKt = (K[j].astype(np.float64), P[j].astype(np.float64), Qi[j].astype(np.float64), M[j].astype(np.float64))
        ref = oracle.solve_batch(p, x0[sel], xref, dtype=np.float32, nthreads=4)
        assert_same(outs[0][0][sel], ref.iter, "iter vs oracle, system %d" % sidx)
        assert_same(outs[0][2][sel], ref.x, "x vs oracle")
        assert_same(outs[0][3][sel], ref.u, "u vs oracle")
        assert_same(outs[0][4][sel], ref.resid, "resid vs oracle")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_per_instance_systems_solve_cartpole_shape(pkg, oracle, dtype):
    """Per-instance systems at 4/1/10 (the cartpole of examples/codegen_cartpole.cpp with perturbed dynamics, input gain and rho):
    batched precompute on the device, solve with instance i on system i, against the oracle system by system (caches as the device
    computed them; work.Q = Q + rho as tiny_codegen stores it)."""
    import copy
    import torch
    sc, m = pkg.problems.cartpole_model()
    base = pkg.problems.cartpole()
    S, per = 16, 40
    rng = np.random.default_rng(7)
    A = np.repeat(m["Adyn"][None], S, 0).copy()
    off = ~np.eye(4, dtype=bool)
    A[:, off] *= (1.0 + 0.1 * rng.uniform(-1, 1, (S, 1)))
    Bm = m["Bdyn"][None] * (1.0 + 0.2 * rng.uniform(-1, 1, (S, 1, 1)))
    Q = np.repeat(m["Q"].reshape(1, -1), S, 0)
    R = np.repeat(m["R"].reshape(1, -1), S, 0)
    rho = float(sc["rho"]) * (1.0 + 0.3 * rng.uniform(-1, 1, S))
    cast = lambda a: a.astype(dtype).astype(np.float64)          # what the device sees
    A, Bm, Q, R, rho = cast(A), cast(Bm), cast(Q), cast(R), cast(rho)
    B = S * per
    idx = np.repeat(np.arange(S), per)
    s = pkg.capi.Solver(base, dtype=dtype, policy="parity")
    sy = pkg.capi.Systems(s, A[idx], Bm[idx], Q[idx], R[idx], rho[idx], q_plus_rho=True)
    x0, xref = pkg.workloads.cartpole_batch(0, B)
    dev = torch.device("cuda:0")
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(dev)
    x = torch.empty((B, 10, 4), dtype=tdt, device=dev); u = torch.empty((B, 9, 1), dtype=tdt, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    sy.solve_raw(f(x0), f(xref), True, x, u, it, st, None)
    torch.cuda.synchronize()
    K, P, Qi, M, Qd = sy.get("Kinf"), sy.get("Pinf"), sy.get("Quu_inv"), sy.get("AmBKt"), sy.get("Q")
    itn, xn, un = it.cpu().numpy(), x.cpu().numpy(), u.cpu().numpy()
    for sidx in range(S):
        j = sidx * per
        p = copy.deepcopy(base)
        p.Adyn, p.Bdyn, p.rho = A[sidx], Bm[sidx], float(rho[sidx])
        p.Q = Qd[j].astype(np.float64)
        p.Kinf, p.Pinf, p.Quu_inv, p.AmBKt = (K[j].astype(np.float64), P[j].astype(np.float64), Qi[j].astype(np.float64), M[j].astype(np.float64))
        ref = oracle.solve_batch(p, x0[j:j + per], xref, dtype=dtype, nthreads=4)
        assert_same(itn[j:j + per], ref.iter, "iter sys %d" % sidx)
        assert_same(xn[j:j + per], ref.x, "x sys %d" % sidx)
        assert_same(un[j:j + per], ref.u, "u sys %d" % sidx)
    assert len(set(itn.tolist())) > 3
