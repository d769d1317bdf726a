"""GPU, at BASELINE.json's FULL size (1,048,576 quadrotor instances, config 2): size-independent properties, plus the
oracle on a 100,000-instance prefix.  Everything through the C ABI with device buffers (as bench.py does)."""
import numpy as np
import pytest

from conftest import assert_same

pytestmark = pytest.mark.gpu
B = 1 << 20


def _solve_device(pkg, solver, x0, xref, lo=0, hi=None):
    import torch
    hi = len(x0) if hi is None else hi
    n = hi - lo
    dev = torch.device("cuda:0")
    x0d = torch.from_numpy(x0[lo:hi]).to(dev)
    xrd = torch.from_numpy(xref).to(dev)
    x = torch.empty((n, 10, 12), device=dev); u = torch.empty((n, 9, 4), device=dev)
    it = torch.empty(n, dtype=torch.int32, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev); rs = torch.empty((n, 4), device=dev)
    solver.solve_raw(n, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs)
    torch.cuda.synchronize()
    return {"x": x, "u": u, "iter": it, "status": st, "resid": rs}


def test_full_size_properties(pkg, oracle, monkeypatch):
    import torch
    prob = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    a = _solve_device(pkg, s, x0, xref)
    assert s.stats()["pattern"] == 1 and s.stats()["iterations"] == int(a["iter"].sum().item())
    # 1. the oracle on a prefix (bit-exact), and the published statistics of the workload (mean 34.2, 7.4 % at max_iter)
    n = 100_000
    ref = oracle.solve_batch(prob, x0[:n], xref, dtype=np.float32, nthreads=16)
    assert_same(a["iter"][:n].cpu().numpy(), ref.iter, "iter prefix")
    assert_same(a["x"][:n].cpu().numpy(), ref.x, "x prefix")
    assert_same(a["u"][:n].cpu().numpy(), ref.u, "u prefix")
    it = a["iter"].cpu().numpy()
    assert abs(it.mean() - 34.1) < 0.2 and abs((it == 100).mean() - 0.072) < 0.005
    # 2. idempotence: the same call again gives the same bits (lane refill order is timing-dependent, results must not be)
    b = _solve_device(pkg, s, x0, xref)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # 3. shard independence: two half batches == the full batch (what the multi-GPU path relies on)
    lo = _solve_device(pkg, s, x0, xref, 0, B // 2)
    hi = _solve_device(pkg, s, x0, xref, B // 2, B)
    for k in a:
        assert torch.equal(a[k], torch.cat([lo[k], hi[k]])), k
    # 4. the dense kernel instance (no structure specialisation) gives the same values everywhere
    monkeypatch.setenv("TMPC_DENSE", "1")
    sd = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    d = _solve_device(pkg, sd, x0, xref)
    assert sd.stats()["pattern"] == 0
    for k in a:
        assert bool((a[k] == d[k]).all()), k          # value equality (+0 == -0)
    # 5. a checksum of checksums over the outputs, stable across runs / kernels (float64 accumulation on the device)
    cs = lambda o: (float(o["x"].double().sum()), float(o["u"].double().sum()), int(o["iter"].sum()))
    assert cs(a) == cs(b) == cs(d)


def test_host_path_with_overlapped_input_copy(pkg, oracle):
    """tmpc_solve(TMPC_MEM_HOST) overlaps the H2D of x0 / per-instance Xref with the kernel (arrival counter + in-kernel
    gate) and the D2H with the kernel (completion counters): 400,000 tracking instances = 4 input chunks, 7 output chunks;
    results must equal the device-buffer path and the oracle."""
    import torch
    prob = pkg.problems.quadrotor(20)
    n = 400_000
    x0, xref = pkg.workloads.quadrotor_tracking_batch(0, n)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    host = s.solve(x0, xref)                                   # pageable numpy buffers
    dev = torch.device("cuda:0")
    x = torch.empty((n, 10, 12), device=dev); u = torch.empty((n, 9, 4), device=dev)
    it = torch.empty(n, dtype=torch.int32, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev); rs = torch.empty((n, 4), device=dev)
    s.solve_raw(n, torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev), False, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs)
    torch.cuda.synchronize()
    assert_same(host["iter"], it.cpu().numpy(), "iter host vs device")
    assert_same(host["x"], x.cpu().numpy(), "x host vs device")
    assert_same(host["u"], u.cpu().numpy(), "u host vs device")
    assert_same(host["resid"], rs.cpu().numpy(), "resid host vs device")
    m = 20_000
    ref = oracle.solve_batch(prob, x0[-m:], xref[-m:], dtype=np.float32, nthreads=16)     # the LAST instances: last input chunk
    assert_same(host["iter"][-m:], ref.iter, "iter tail vs oracle")
    assert_same(host["u"][-m:], ref.u, "u tail vs oracle")
    # pinned buffers through the same path
    px0 = torch.from_numpy(x0).pin_memory(); pxr = torch.from_numpy(xref).pin_memory()
    hx = torch.empty((n, 10, 12)).pin_memory(); hu = torch.empty((n, 9, 4)).pin_memory()
    hit = torch.empty(n, dtype=torch.int32).pin_memory(); hst = torch.empty(n, dtype=torch.int32).pin_memory(); hrs = torch.empty((n, 4)).pin_memory()
    s.solve_raw(n, px0, pxr, False, pkg.capi.TMPC_MEM_HOST, hx, hu, hit, hst, hrs)
    assert_same(hit.numpy(), host["iter"], "iter pinned")
    assert_same(hx.numpy(), host["x"], "x pinned")


def _device_solve_any(pkg, prob, x0, xref, warm=None, solver=None):
    import torch
    dev = torch.device("cuda:0")
    n = len(x0)
    s = solver or pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    x = torch.empty((n, prob.N, prob.nx), device=dev); u = torch.empty((n, prob.N - 1, prob.nu), device=dev)
    it = torch.empty(n, dtype=torch.int32, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev); rs = torch.empty((n, 4), device=dev)
    s.solve_raw(n, torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev), xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, warm=warm)
    torch.cuda.synchronize()
    return s, {"x": x, "u": u, "iter": it, "status": st, "resid": rs}


def _check_prefix(out, ref, n, what):
    for k in ("iter", "status", "x", "u", "resid"):
        assert_same(out[k][:n].cpu().numpy(), getattr(ref, k), "%s %s" % (what, k))


def test_config3_tracking_full_size(pkg, oracle):
    """BASELINE configs[2] at its named size: 4,194,304 tracking instances with per-instance reference windows.  Oracle on a
    20,000-instance prefix and on the last 10,000; window periodicity (instances b and b + 290 k share k_b but not x0, so only the
    statistics repeat); idempotence."""
    import torch
    prob = pkg.problems.quadrotor(20)
    B3 = 1 << 22
    x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B3)
    s, a = _device_solve_any(pkg, prob, x0, xref)
    assert s.stats()["instances"] == B3 and s.stats()["iterations"] == int(a["iter"].sum().item())
    n = 20_000
    _check_prefix(a, oracle.solve_batch(prob, x0[:n], xref[:n], dtype=np.float32, nthreads=16), n, "config 3 prefix")
    m = 10_000
    ref = oracle.solve_batch(prob, x0[-m:], xref[-m:], dtype=np.float32, nthreads=16)
    assert_same(a["iter"][-m:].cpu().numpy(), ref.iter, "config 3 tail iter")
    assert_same(a["u"][-m:].cpu().numpy(), ref.u, "config 3 tail u")
    assert bool((a["status"] == 1).all())                       # every tracking instance converges (10-15 iterations)
    _, b = _device_solve_any(pkg, prob, x0, xref, solver=s)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_config4_cartpole_full_size(pkg, oracle):
    """BASELINE configs[3] at its named size: 16,777,216 cartpole instances.  Oracle on a 50,000-instance prefix, shard
    independence (four quarter batches == the full batch), statistics of the seeded workload."""
    import torch
    prob = pkg.problems.cartpole()
    B4 = 1 << 24
    x0, xref = pkg.workloads.cartpole_batch(0, B4)
    s, a = _device_solve_any(pkg, prob, x0, xref)
    assert s.stats()["instances"] == B4 and s.stats()["iterations"] == int(a["iter"].sum().item())
    n = 50_000
    _check_prefix(a, oracle.solve_batch(prob, x0[:n], xref, dtype=np.float32, nthreads=16), n, "config 4 prefix")
    q = B4 // 4
    for j in (1, 3):
        _, part = _device_solve_any(pkg, prob, x0[j * q:(j + 1) * q], xref, solver=s)
        for k in a:
            assert torch.equal(a[k][j * q:(j + 1) * q], part[k]), (j, k)
    it = a["iter"].cpu().numpy()
    assert abs(it.mean() - 55.9) < 0.5


def test_config5_large_warm_full_size(pkg, oracle):
    """BASELINE configs[4] at its named size: 262,144 instances of the 32/8/50 system, cold solve, x0 perturbed by 1 %, re-solve
    warm-started from the state left in HBM.  Oracle (cold and warm, state included) on a 2,000-instance prefix; the cold
    solve's statistics."""
    import torch
    prob = pkg.problems.random_system()
    B5 = 1 << 18
    dev = torch.device("cuda:0")
    x0, xref = pkg.workloads.random_system_batch(0, B5)
    warm = {k: torch.zeros((B5, 49, 8) if k in "dyz" else (B5, 50, 32), device=dev) for k in ("d", "y", "g", "v", "z")}
    s, a = _device_solve_any(pkg, prob, x0, xref, warm=warm)
    n = 2_000
    r1 = oracle.solve_batch(prob, x0[:n], xref, dtype=np.float32, want_state=True, nthreads=16)
    _check_prefix(a, r1, n, "config 5 cold prefix")
    for k in warm:
        assert_same(warm[k][:n].cpu().numpy(), r1.state[k], "config 5 cold state " + k)
    it = a["iter"].cpu().numpy()
    assert abs(it.mean() - 71) < 2
    x1 = pkg.workloads.perturb_x0(x0, 0)
    _, b = _device_solve_any(pkg, prob, x1, xref, warm=warm, solver=s)
    r2 = oracle.solve_batch(prob, x1[:n], xref, dtype=np.float32, warm={k: r1.state[k] for k in ("d", "y", "g", "v", "z")}, want_state=True,
                            nthreads=16)
    _check_prefix(b, r2, n, "config 5 warm prefix")
    for k in warm:
        assert_same(warm[k][:n].cpu().numpy(), r2.state[k], "config 5 warm state " + k)
    assert b["iter"].float().mean().item() < 0.6 * it.mean()      # the warm start pays (mean 27 vs 71 iterations)
