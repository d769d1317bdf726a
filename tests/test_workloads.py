"""Seeded generator + problem data: deterministic, shardable, and consistent with the reference's data."""
import numpy as np


def test_splitmix_known_values(pkg):
    W = pkg.workloads
    # splitmix64 reference values (seed 0 stream: first outputs of the standard generator)
    assert int(W.splitmix64(np.uint64(0))) == 0xE220A8397B1DCDAF
    assert int(W.splitmix64(np.uint64(0x9E3779B97F4A7C15))) == 0x6E789E6AA1B965F4
    u = W.u01(1234, np.arange(1000, dtype=np.uint64))
    assert (u >= 0).all() and (u < 1).all() and np.all(u == u.astype(np.float32))


def test_batches_are_shardable(pkg):
    W = pkg.workloads
    full, xr = W.quadrotor_hover_batch(0, 1000, mult=0.25)
    parts = [W.quadrotor_hover_batch(a, b, mult=0.25)[0] for a, b in ((0, 250), (250, 700), (700, 1000))]
    assert np.array_equal(np.concatenate(parts), full)
    x0, xref = W.quadrotor_tracking_batch(285, 295)
    table = pkg.problems.quadrotor_trajectory()
    assert np.array_equal(xref[5], table[:, 0:10].T.astype(np.float32))  # b = 290 -> window 0
    assert xref.shape == (10, 10, 12)


def test_shipped_cache_matches_precompute(pkg):
    """the shipped caches are the codegen.cpp:254-292 recursion on Q+rho, R+rho."""
    P = pkg.problems
    for hz, tol in ((20, 5e-5), (50, 1e-4), (100, 2e-4)):
        q = P.quadrotor(hz)
        c = P.precompute_cache(q.Adyn, q.Bdyn, q.Q, q.R, q.rho)
        assert np.abs(c["Kinf"] - q.Kinf).max() < tol
        assert np.abs(c["Quu_inv"] - q.Quu_inv).max() < 1e-6
        assert np.abs(c["AmBKt"] - q.AmBKt).max() < 2e-3   # (A - B K)^T amplifies dK by |B| ~ 10
