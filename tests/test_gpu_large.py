"""GPU parity for the large shape (BASELINE config 5: nx=32, nu=8, N=50, warp-per-instance kernel) through the
C ABI against the CPU oracle: cold solve, warm-state write-back, warm-started re-solve -- all bit-exact in fp32
(PARITY policy); FAST policy within 1e-4 on equal-iteration instances."""
import copy

import numpy as np
import pytest

from conftest import assert_same

pytestmark = pytest.mark.gpu

WARM_KEYS = ("d", "y", "g", "v", "z")


def _zeros_warm(prob, B):
    return {k: np.zeros((B, prob.N - 1, prob.nu) if k in "dyz" else (B, prob.N, prob.nx), np.float32) for k in WARM_KEYS}


def test_large_cold_bit_exact(pkg, oracle):
    prob = pkg.problems.random_system()
    B = 300
    x0, xref = pkg.workloads.random_system_batch(0, B)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    out = s.solve(x0, xref)
    assert (ref.status == 1).any() and (ref.status == 11).any()
    for k in ("iter", "status", "resid", "x", "u"):
        assert_same(out[k], getattr(ref, k), k)
    st = s.stats()
    assert st["iterations"] == int(ref.iter.sum()) and st["solved"] == int((ref.status == 1).sum())
    assert st["parity_pinned"] == 1 and st["launches"] == 1


def test_large_warm_start_config5_sequence(pkg, oracle):
    """Config 5 as BASELINE states it: cold solve, x0 perturbed by 1 %, re-solve with {d,y,g,v,z} carried in HBM."""
    prob = pkg.problems.random_system()
    B = 200
    x0, xref = pkg.workloads.random_system_batch(0, B)
    r1 = oracle.solve_batch(prob, x0, xref, dtype=np.float32, want_state=True, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    o1 = s.solve(x0, xref, warm=_zeros_warm(prob, B))
    assert_same(o1["iter"], r1.iter, "iter")
    assert_same(o1["x"], r1.x, "x")
    for k in WARM_KEYS:
        assert_same(o1["warm"][k], r1.state[k], "warm." + k)
    x1 = pkg.workloads.perturb_x0(x0, 0)
    r2 = oracle.solve_batch(prob, x1, xref, dtype=np.float32, warm={k: r1.state[k] for k in WARM_KEYS},
                            want_state=True, nthreads=8)
    o2 = s.solve(x1, xref, warm=o1["warm"])
    assert r2.iter.mean() < r1.iter.mean()
    for k in ("iter", "status", "resid", "x", "u"):
        assert_same(o2[k], getattr(r2, k), k + " (warm)")
    for k in WARM_KEYS:
        assert_same(o2["warm"][k], r2.state[k], "warm2." + k)


def test_large_settings_and_per_instance_xref(pkg, oracle):
    prob = copy.deepcopy(pkg.problems.random_system())
    prob.check_termination = 4
    prob.max_iter = 37
    prob.en_state_bound = 0
    B = 97   # ragged: not a multiple of the 12 instances per block
    x0, _ = pkg.workloads.random_system_batch(0, B, amp=0.7)
    xref = (0.05 * pkg.workloads._noise(99, 0, B, prob.N * prob.nx)).astype(np.float32).reshape(B, prob.N, prob.nx)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, want_state=True, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    out = s.solve(x0, xref, warm=_zeros_warm(prob, B))
    for k in ("iter", "status", "resid", "x", "u"):
        assert_same(out[k], getattr(ref, k), k)
    for k in WARM_KEYS:
        assert_same(out["warm"][k], ref.state[k], "warm." + k)
    # B = 1 and the empty batch
    o1 = s.solve(x0[:1], xref[:1])
    assert_same(o1["x"], ref.x[:1], "x B=1")
    assert s.solve(np.zeros((0, 32), np.float32), xref[:0].reshape(0, prob.N, prob.nx) if False else np.zeros((prob.N, 32), np.float32))["iter"].shape == (0,)


def test_large_fast_policy(pkg, oracle):
    prob = pkg.problems.random_system()
    B = 300
    x0, xref = pkg.workloads.random_system_batch(0, B, amp=0.5)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)
    out = pkg.capi.Solver(prob, dtype=np.float32, policy="fast").solve(x0, xref)
    same = out["iter"] == ref.iter
    assert (~same).mean() <= 0.10, "iteration mismatch rate %.4f" % (~same).mean()
    scale = np.maximum(np.abs(ref.x[same]).max(), 1.0)
    assert np.abs(out["x"][same] - ref.x[same]).max() / scale <= 1e-4
    assert np.abs(out["u"][same] - ref.u[same]).max() <= 1e-4


def test_large_step_functions(pkg, oracle):
    """The six reference step functions at 32/8/50 (thread-per-instance step kernels) against the oracle."""
    from oracle.pyoracle import ws_size
    prob = pkg.problems.random_system()
    nx, nu, N = prob.nx, prob.nu, prob.N
    rng = np.random.default_rng(7)
    B = 8
    names = [("x", N * nx), ("u", (N - 1) * nu), ("q", N * nx), ("r", (N - 1) * nu), ("p", N * nx), ("d", (N - 1) * nu),
             ("v", N * nx), ("vnew", N * nx), ("z", (N - 1) * nu), ("znew", (N - 1) * nu), ("g", N * nx),
             ("y", (N - 1) * nu), ("Xref", N * nx), ("resid", 4)]
    assert sum(n for _, n in names) == ws_size(nx, nu, N)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    for which in range(6):
        img = rng.uniform(-1, 1, size=(B, ws_size(nx, nu, N))).astype(np.float32)
        ws, off = {}, 0
        for k, n in names:
            shape = (B, 4) if k == "resid" else ((B, N - 1, nu) if n == (N - 1) * nu else (B, N, nx))
            ws[k] = np.ascontiguousarray(img[:, off:off + n]).reshape(shape)
            off += n
        term = s.step(which, ws, it=2)
        for b in range(B):
            rc, exp = oracle.step(prob, which, img[b], it=2)
            off = 0
            for k, n in names:
                if k != "Xref":
                    assert_same(ws[k][b].reshape(-1), exp[off:off + n], "step %d %s" % (which, k))
                off += n
            if which == 4:
                assert term[b] == rc


@pytest.mark.parametrize("variant", ["", "warp4", "warp4x2", "warp_smem"])
def test_large_kernel_variants_and_slot_refill(pkg, oracle, monkeypatch, variant):
    """Default = one instance per warp, g / v in tensor memory (tmpc_kernel_warp.cuh); TMPC_KERNEL=warp4 = four instances per warp
    (tmpc_kernel_warp4.cuh), warp4x2 = the same with eight warps per SM (g / v of two slots per warp in L2-resident scratch),
    warp_smem = one instance per warp with all state in shared memory.  1,100 instances on a 3-block grid cap would be too few to refill slots, so the batch is several times the resident
    slots of the test GPU only in the sense that every slot is refilled many times in index order (148 SMs x 16 = 2,368 resident:
    use a ragged 5,003 with a short max_iter so that the oracle stays cheap), cold and warm, state included."""
    if variant:
        monkeypatch.setenv("TMPC_KERNEL", variant)
    prob = copy.deepcopy(pkg.problems.random_system())
    prob.max_iter = 12
    B = 5003
    x0, xref = pkg.workloads.random_system_batch(0, B, amp=0.3)
    r1 = oracle.solve_batch(prob, x0, xref, dtype=np.float32, want_state=True, nthreads=8)
    assert (r1.status == 1).any() and (r1.status == 11).any()
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    cold = s.solve(x0, xref)
    for k in ("iter", "status", "resid", "x", "u"):
        assert_same(cold[k], getattr(r1, k), "%s cold %s" % (variant, k))
    o1 = s.solve(x0, xref, warm=_zeros_warm(prob, B))
    for k in WARM_KEYS:
        assert_same(o1["warm"][k], r1.state[k], "%s warm.%s" % (variant, k))
    x1 = pkg.workloads.perturb_x0(x0, 0)
    r2 = oracle.solve_batch(prob, x1, xref, dtype=np.float32, warm={k: r1.state[k] for k in WARM_KEYS}, want_state=True, nthreads=8)
    o2 = s.solve(x1, xref, warm=o1["warm"], outputs=("x", "u", "u0", "iter", "status", "resid"))
    for k in ("iter", "status", "resid", "x", "u"):
        assert_same(o2[k], getattr(r2, k), "%s re-solve %s" % (variant, k))
    assert_same(o2["u0"], r2.u[:, 0, :], "%s re-solve u0" % variant)
    for k in WARM_KEYS:
        assert_same(o2["warm"][k], r2.state[k], "%s warm2.%s" % (variant, k))
