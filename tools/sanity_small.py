"""Tiny invocation of every kernel family (for compute-sanitizer runs): each result is also checked against the oracle."""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import OracleLib  # noqa: E402

pkg = load_package()
capi = pkg.capi
ora = OracleLib()


def check(name, out, ref):
    ok = all(np.array_equal(out[k], getattr(ref, k)) for k in ("iter", "x", "u"))
    print("%-34s %s" % (name, "ok" if ok else "MISMATCH"), flush=True)
    assert ok


def zeros_warm(p, B, dt):
    return {k: np.zeros((B, p.N - 1, p.nu) if k in "dyz" else (B, p.N, p.nx), dt) for k in ("d", "y", "g", "v", "z")}


quad = copy.deepcopy(pkg.problems.quadrotor(20)); quad.max_iter = 12
x0, xref = pkg.workloads.quadrotor_hover_batch(0, 300, mult=0.25)
for env in ({}, {"TMPC_DENSE": "1"}):
    os.environ.update(env)
    for dt in (np.float32, np.float64):
        s = capi.Solver(quad, dtype=dt, policy="parity")
        check("quad %s %s cold" % (np.dtype(dt).name, env), s.solve(x0, xref), ora.solve_batch(quad, x0, xref, dtype=dt, nthreads=4))
        check("quad %s %s warm" % (np.dtype(dt).name, env), s.solve(x0, xref, warm=zeros_warm(quad, 300, dt)), ora.solve_batch(quad, x0, xref, dtype=dt, nthreads=4))
        s.close()
    for k in env:
        del os.environ[k]
s = capi.Solver(quad, dtype=np.float32, policy="fast"); s.solve(x0, xref); s.close(); print("quad f32 fast ran", flush=True)
cart = copy.deepcopy(pkg.problems.cartpole()); cart.max_iter = 12
cx0, cxr = pkg.workloads.cartpole_batch(0, 700)
s = capi.Solver(cart, dtype=np.float32, policy="parity"); check("cartpole f32", s.solve(cx0, cxr), ora.solve_batch(cart, cx0, cxr, dtype=np.float32, nthreads=4)); s.close()
big = copy.deepcopy(pkg.problems.random_system()); big.max_iter = 6
bx0, bxr = pkg.workloads.random_system_batch(0, 40)
for env in ({}, {"TMPC_KERNEL": "warp_smem"}):
    os.environ.update(env)
    s = capi.Solver(big, dtype=np.float32, policy="parity")
    check("large 32/8/50 %s cold" % env, s.solve(bx0, bxr), ora.solve_batch(big, bx0, bxr, dtype=np.float32, nthreads=4))
    check("large 32/8/50 %s warm" % env, s.solve(bx0, bxr, warm=zeros_warm(big, 40, np.float32)), ora.solve_batch(big, bx0, bxr, dtype=np.float32, nthreads=4))
    s.close()
    for k in env:
        del os.environ[k]
# batch API + rollout
s = capi.Solver(quad, dtype=np.float32, policy="parity")
b = capi.Batch(s, 200); b.set_x0(x0[:200]); b.set_xref_table(pkg.problems.quadrotor_trajectory().T, np.arange(200) % 250)
h = b.rollout(3); print("rollout ran, iters", int(h["iter"].sum()), flush=True); b.close()
# systems: precompute + both solve kernels
S = 150
rng = np.random.default_rng(0)
A = np.repeat(quad.Adyn[None], S, 0) * (1.0 + 0.01 * rng.uniform(-1, 1, (S, 1, 1))); Bm = np.repeat(quad.Bdyn[None], S, 0)
Q = np.repeat(quad.Q[None], S, 0); R = np.repeat(quad.R[None], S, 0); rho = np.full(S, 5.0)
dev = torch.device("cuda:0")
for env in ({}, {"TMPC_KERNEL": "sys_global"}):
    os.environ.update(env)
    sy = capi.Systems(s, A, Bm, Q, R, rho)
    x = torch.empty((S, 10, 12), device=dev); u = torch.empty((S, 9, 4), device=dev); it = torch.empty(S, dtype=torch.int32, device=dev)
    st = torch.empty(S, dtype=torch.int32, device=dev)
    sy.solve_raw(torch.from_numpy(x0[:S]).to(dev), torch.from_numpy(xref).to(dev), True, x, u, it, st, None)
    torch.cuda.synchronize()
    print("systems %s ran, sweeps %d..%d, iters %d" % (env, sy.get("sweeps").min(), sy.get("sweeps").max(), int(it.sum().item())), flush=True)
    sy.close()
    for k in env:
        del os.environ[k]
print("ALL OK", flush=True)
