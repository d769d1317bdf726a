"""Would a tensor-core (TF32) evaluation of the batched mat-vecs meet north_star's accuracy bar?  CPU experiment (numpy):
the ADMM loop of tiny_solve (admm.cpp:111-152) on a batch of quadrotor instances in float32, once with IEEE float32
products (the FAST policy's arithmetic up to summation order) and once with both operands of every mat-vec product
rounded to TF32 (10-bit mantissa, what tcgen05.mma kind::tf32 consumes; accumulation stays fp32), and the 3xTF32 split
(a_hi b_hi + a_hi b_lo + a_lo b_hi).  Reports, against the bit-exact oracle: share of instances whose iteration count
differs, and max relative x/u error on the instances whose count agrees (north_star: counts must match, 1e-4 relative)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import OracleLib  # noqa: E402  (tools/ may use the checker; nothing here ships)

pkg = load_package()
f32 = np.float32


def tf32(a):
    """round-to-nearest-even to a 10-bit mantissa"""
    b = np.ascontiguousarray(a, f32).view(np.uint32).astype(np.uint64)
    b = (b + 0xFFF + ((b >> 13) & 1)) & ~np.uint64(0x1FFF)
    return b.astype(np.uint32).view(f32)


def mm(M, X, mode):
    """X [B, k] times M^T [k, r] -> [B, r] with the chosen product arithmetic, fp32 accumulation"""
    M = M.astype(f32); X = X.astype(f32)
    if mode == "fp32":
        return (X[:, None, :] * M[None, :, :]).sum(axis=2, dtype=f32)
    Mh, Xh = tf32(M), tf32(X)
    if mode == "tf32":
        return (Xh[:, None, :] * Mh[None, :, :]).sum(axis=2, dtype=f32)
    Ml, Xl = tf32(M - Mh), tf32(X - Xh)
    return ((Xh[:, None, :] * Mh[None, :, :]) + (Xh[:, None, :] * Ml[None, :, :]) + (Xl[:, None, :] * Mh[None, :, :])).sum(axis=2, dtype=f32)


def solve(prob, x0, xref, mode):
    B = x0.shape[0]; nx, nu, N = prob.nx, prob.nu, prob.N
    K, A, Bm, Qi, M, P = (np.asarray(m, f32) for m in (prob.Kinf, prob.Adyn, prob.Bdyn, prob.Quu_inv, prob.AmBKt, prob.Pinf))
    rho = f32(prob.rho)
    x = np.zeros((B, N, nx), f32); u = np.zeros((B, N - 1, nu), f32)
    d = np.zeros_like(u); y = np.zeros_like(u); z = np.zeros_like(u); g = np.zeros_like(x); v = np.zeros_like(x)
    x[:, 0] = x0
    it = np.zeros(B, np.int32); done = np.zeros(B, bool)
    xr = np.broadcast_to(xref.astype(f32), (B, N, nx))
    pN = -mm(P.T, xr[:, N - 1], mode)
    xo = np.zeros_like(x); uo = np.zeros_like(u)
    for k in range(prob.max_iter):
        for i in range(N - 1):
            u[:, i] = -mm(K, x[:, i], mode) - d[:, i]
            x[:, i + 1] = mm(A, x[:, i], mode) + mm(Bm, u[:, i], mode)
        zn = np.clip(u + y, prob.u_min.astype(f32), prob.u_max.astype(f32)); vn = np.clip(x + g, prob.x_min.astype(f32), prob.x_max.astype(f32))
        y = (y + u) - zn; g = (g + x) - vn
        r = -rho * (zn - y); q = -(xr * prob.Q.astype(f32)) - rho * (vn - g)
        p = pN - rho * (vn[:, N - 1] - g[:, N - 1])
        conv = ((np.abs(x - vn).max(axis=(1, 2)) < prob.abs_pri_tol) & (np.abs(u - zn).max(axis=(1, 2)) < prob.abs_pri_tol) &
                (rho * np.abs(v - vn).max(axis=(1, 2)) < prob.abs_dua_tol) & (rho * np.abs(z - zn).max(axis=(1, 2)) < prob.abs_dua_tol))
        newly = ~done & (conv | (k == prob.max_iter - 1))
        it[newly] = k + 1; xo[newly] = x[newly]; uo[newly] = u[newly]
        done |= newly
        if done.all():
            break
        v, z = vn, zn
        for i in range(N - 2, -1, -1):
            d[:, i] = mm(Qi, mm(Bm.T, p, mode) + r[:, i], mode)
            p = q[:, i] + mm(M, p, mode) - mm(K.T, r[:, i], mode)
    return it, xo, uo


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    prob = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    ref = OracleLib().solve_batch(prob, x0, xref, dtype=f32, nthreads=8)
    for mode in ("fp32", "tf32", "3xtf32"):
        it, x, u = solve(prob, x0, xref, mode)
        same = it == ref.iter
        sx = np.abs(ref.x[same]).max(axis=(1, 2), keepdims=True) + 1e-6; su = np.abs(ref.u[same]).max(axis=(1, 2), keepdims=True) + 1e-6
        print("%-7s iteration count differs on %5.2f %% of %d instances (max |d iter| %d); on the others max rel err x %.2e, u %.2e" %
              (mode, 100 * (1 - same.mean()), B, np.abs(it - ref.iter).max(), (np.abs(x[same] - ref.x[same]) / sx).max(),
               (np.abs(u[same] - ref.u[same]) / su).max()))
