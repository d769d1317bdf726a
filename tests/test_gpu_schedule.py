"""The longest-expected-first schedule (tmpc_api.cu lpt_prepare; SolveArgs::order) forced on (TMPC_LPT=1) for EVERY kernel
family at small batch sizes: outputs are indexed by instance and must equal the oracle's, instance by instance, whatever
order the lanes claimed the work in."""
import copy

import numpy as np
import pytest

from conftest import assert_same

pytestmark = pytest.mark.gpu


def _solve_device(pkg, prob, x0, xref, dtype, warm=None, policy="parity"):
    import torch
    dev = torch.device("cuda:0")
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    B = x0.shape[0]
    s = pkg.capi.Solver(prob, dtype=dtype, policy=policy)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(dev)
    x = torch.empty((B, prob.N, prob.nx), dtype=tdt, device=dev); u = torch.empty((B, prob.N - 1, prob.nu), dtype=tdt, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), dtype=tdt, device=dev)
    w = None if warm is None else {k: f(v) for k, v in warm.items()}
    s.solve_raw(B, f(x0), f(xref), xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, warm=w,
                stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = {"x": x.cpu().numpy(), "u": u.cpu().numpy(), "iter": it.cpu().numpy(), "status": st.cpu().numpy(), "resid": rs.cpu().numpy()}
    if w is not None:
        out["warm"] = {k: v.cpu().numpy() for k, v in w.items()}
    return s.stats(), out


CASES = ["quadrotor_f32", "quadrotor_f32_dense", "quadrotor_f64", "cartpole_f32", "cartpole_f64", "large_cold", "large_warm", "generic_6_3_20",
         "generic_9_4_7_f64"]


@pytest.mark.parametrize("case", CASES)
def test_forced_schedule_matches_oracle(pkg, oracle, case, monkeypatch):
    monkeypatch.setenv("TMPC_LPT", "1")
    dtype = np.float64 if case.endswith("f64") else np.float32
    warm = None
    if case.startswith("quadrotor"):
        prob = pkg.problems.quadrotor(20)
        x0, xref = pkg.workloads.quadrotor_hover_batch(0, 2777, mult=0.5)
        if case.endswith("dense"):
            monkeypatch.setenv("TMPC_DENSE", "1")
    elif case.startswith("cartpole"):
        prob = pkg.problems.cartpole()
        x0, xref = pkg.workloads.cartpole_batch(0, 4001)
    elif case.startswith("large"):
        prob = pkg.problems.random_system()
        x0, xref = pkg.workloads.random_system_batch(0, 75, amp=0.6)
        if case == "large_warm":
            r0 = oracle.solve_batch(prob, x0, xref, dtype=dtype, want_state=True, nthreads=8)
            warm = {k: r0.state[k] for k in ("d", "y", "g", "v", "z")}
            x0 = pkg.workloads.perturb_x0(x0, 0)
    else:
        nx, nu, N = (6, 3, 20) if "6_3_20" in case else (9, 4, 7)
        prob = pkg.problems.random_system(nx, nu, N, seed=7 + nx)
        rng = np.random.default_rng(nx)
        x0 = rng.uniform(-2, 2, (1500, nx)).astype(np.float32)
        x0[::2] *= np.float32(0.1)
        xref = np.zeros((N, nx), np.float32)
    ref = oracle.solve_batch(prob, x0, xref, dtype=dtype, warm=copy.deepcopy(warm), want_state=True, nthreads=8)
    st, out = _solve_device(pkg, prob, x0, xref, dtype, warm=warm)
    assert st["scheduled"] == 1, "the schedule was not used"
    for k in ("iter", "status", "x", "u", "resid"):
        assert_same(out[k], getattr(ref, k), case + " " + k)
    if warm is not None:
        for k in warm:
            assert_same(out["warm"][k], ref.state[k], case + " warm " + k)
    assert st["iterations"] == int(ref.iter.sum()) and st["instances"] == x0.shape[0]


def test_forced_schedule_per_instance_systems(pkg, oracle, monkeypatch):
    import torch
    from test_gpu_systems import _systems
    monkeypatch.setenv("TMPC_LPT", "1")
    S, per = 16, 40
    base, A, Bm, Q, R, rho = _systems(pkg, S, np.float32)
    B = S * per
    idx = np.repeat(np.arange(S), per)
    s = pkg.capi.Solver(base, dtype=np.float32, policy="parity")
    sy = pkg.capi.Systems(s, A[idx], Bm[idx], Q[idx], R[idx], rho[idx])
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.4)
    dev = torch.device("cuda:0")
    x = torch.empty((B, 10, 12), device=dev); u = torch.empty((B, 9, 4), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    sy.solve_raw(torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev), True, x, u, it, st, None)
    torch.cuda.synchronize()
    assert s.stats()["scheduled"] == 1
    K, P, Qi, M = sy.get("Kinf"), sy.get("Pinf"), sy.get("Quu_inv"), sy.get("AmBKt")
    itn, xn, un = it.cpu().numpy(), x.cpu().numpy(), u.cpu().numpy()
    for sidx in range(S):
        p = copy.deepcopy(base)
        j = sidx * per
        p.Adyn, p.Bdyn, p.Q, p.rho = A[sidx].astype(np.float64), Bm[sidx].astype(np.float64), Q[sidx].astype(np.float64), float(rho[sidx])
        p.Kinf, p.Pinf, p.Quu_inv, p.AmBKt = (K[j].astype(np.float64), P[j].astype(np.float64), Qi[j].astype(np.float64), M[j].astype(np.float64))
        ref = oracle.solve_batch(p, x0[j:j + per], xref, dtype=np.float32, nthreads=4)
        assert_same(itn[j:j + per], ref.iter, "iter sys %d" % sidx)
        assert_same(xn[j:j + per], ref.x, "x sys %d" % sidx)
        assert_same(un[j:j + per], ref.u, "u sys %d" % sidx)
