"""Output mask (u0-only mode), the host pipeline's tail-sorted schedule and its fall-backs, launch ordering on one ctx, and
the one-process multi-device solve (tmpc_multi) -- all through the C ABI, against the CPU oracle, bit for bit.

u0 = u(:,0), the control an MPC loop applies (quadrotor_hovering.cpp:110).  A controls-only call (x = u = NULL) takes the
kernel's no-emission path, so it is checked on every kernel family."""
import os

import numpy as np
import pytest

from conftest import assert_same

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


CASES = [("quadrotor", np.float32, 5000), ("quadrotor", np.float64, 3000), ("cartpole", np.float32, 6000),
         ("cartpole", np.float64, 2000), ("large", np.float32, 200), ("generic", np.float32, 500)]


def _case(pkg, name, B):
    P, W = pkg.problems, pkg.workloads
    if name == "quadrotor":
        prob = P.quadrotor(20)
        x0, xref = W.quadrotor_hover_batch(0, B, mult=0.25)
    elif name == "cartpole":
        prob = P.cartpole()
        x0, xref = W.cartpole_batch(0, B)
    elif name == "large":
        prob = P.random_system()
        x0, xref = W.random_system_batch(0, B)
    else:   # a shape without a specialised kernel: the run-time-shape kernel
        prob = P.random_system(nx=6, nu=3, N=8)
        x0, xref = W.random_system_batch(0, B, N=8, nx=6)
    return prob, x0, xref


@pytest.mark.parametrize("name,dtype,B", CASES)
def test_u0_only_and_u0_with_full_outputs(pkg, oracle, name, dtype, B):
    prob, x0, xref = _case(pkg, name, B)
    ref = oracle.solve_batch(prob, x0, xref, dtype=dtype, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")
    only = s.solve(x0, xref, outputs=("u0", "iter", "status"))          # controls-only: no emission pass
    assert_same(only["iter"], ref.iter, "iter")
    assert_same(only["status"], ref.status, "status")
    assert_same(only["u0"], ref.u[:, 0, :], "u0")
    assert s.stats()["iterations"] == int(ref.iter.sum())
    full = s.solve(x0, xref, outputs=("x", "u", "u0", "iter", "status", "resid"))
    assert_same(full["u0"], ref.u[:, 0, :], "u0 (with x,u)")
    assert_same(full["x"], ref.x, "x")
    assert_same(full["u"], ref.u, "u")
    assert_same(full["resid"], ref.resid, "resid")


def test_u0_only_device_memory(pkg, oracle):
    """Device-resident controls-only solve (no emission pass): same numbers."""
    torch = _torch()
    prob = pkg.problems.quadrotor(20)
    B = 100_000
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)
    dev = torch.device("cuda:0")
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
    u0 = torch.empty((B, 4), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    s.solve_raw(B, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, None, None, it, st, None, u0=u0)
    torch.cuda.synchronize()
    trips_u0 = s.stats()["trips"]
    assert_same(it.cpu().numpy(), ref.iter, "iter")
    assert_same(u0.cpu().numpy(), ref.u[:, 0, :], "u0")
    x = torch.empty((B, 10, 12), device=dev)
    u = torch.empty((B, 9, 4), device=dev)
    s.solve_raw(B, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, None)
    torch.cuda.synchronize()
    assert_same(u.cpu().numpy(), ref.u, "u")
    # no emission or speculative trips in the controls-only solve: lane-trips = iterations + idle lanes at the tail only
    assert trips_u0 >= int(ref.iter.sum())


def test_u0_warm_start(pkg, oracle):
    """Warm-started solves still emit (g / y write-back rides on the emission sweep): u0 comes out of that sweep."""
    prob = pkg.problems.quadrotor(20)
    B = 2000
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    ref1 = oracle.solve_batch(prob, x0, xref, dtype=np.float32, want_state=True, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    z = lambda *sh: np.zeros(sh, np.float32)
    warm = {"d": z(B, 9, 4), "y": z(B, 9, 4), "z": z(B, 9, 4), "g": z(B, 10, 12), "v": z(B, 10, 12)}
    o1 = s.solve(x0, xref, warm=warm, outputs=("u0", "iter", "status"))
    assert_same(o1["u0"], ref1.u[:, 0, :], "u0 warm #1")
    assert_same(o1["iter"], ref1.iter, "iter warm #1")
    for k in ("d", "y", "z", "g", "v"):
        assert_same(o1["warm"][k], ref1.state[k], "warm " + k)


def test_host_pipeline_tail_sorted_schedule(pkg, oracle):
    """A host batch large enough for the tail-sorted schedule (leading quarter ranked and claimed last) gives the same
    per-instance results; so do the fall-backs without stream memory operations and without H2D overlap."""
    prob = pkg.problems.quadrotor(20)
    B = 120_000
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)

    def check(env, scheduled):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
            for outputs in (("x", "u", "iter", "status", "resid"), ("u0", "iter", "status")):
                out = s.solve(x0, xref, outputs=outputs)
                assert_same(out["iter"], ref.iter, "iter %r" % (env,))
                assert_same(out["status"], ref.status, "status")
                if "x" in out:
                    assert_same(out["x"], ref.x, "x")
                    assert_same(out["u"], ref.u, "u")
                    assert_same(out["resid"], ref.resid, "resid")
                else:
                    assert_same(out["u0"], ref.u[:, 0, :], "u0")
                assert s.stats()["scheduled"] == scheduled, (env, s.stats())
            s.close()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    check({}, 2)
    check({"TMPC_LPT": "0"}, 0)
    # as on a driver without cuStreamWaitValue32 / cuStreamWriteValue32: inputs complete before the launch, outputs copied
    # after it -> nothing needs index order, the full longest-expected-first schedule applies
    check({"TMPC_NO_STREAM_MEMOPS": "1"}, 1)
    check({"TMPC_NO_H2D_OVERLAP": "1"}, 0)


def test_input_gate_timeout_is_an_error_not_a_hang(pkg):
    """If an input chunk's arrival is never announced the lanes give up after ~2 s; the host must come back with an error
    (the completion counters are released after the kernel), not wait forever."""
    prob = pkg.problems.quadrotor(20)
    B = 150_000
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.1)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    os.environ["TMPC_TEST_GATE_STALL"] = "1"
    try:
        with pytest.raises(pkg.capi.TmpcError, match="input gate timed out"):
            s.solve(x0, xref)
    finally:
        del os.environ["TMPC_TEST_GATE_STALL"]
    out = s.solve(x0[:4000], xref)                # the ctx is still usable
    assert (out["iter"] > 0).all()


def test_two_streams_on_one_ctx_are_serialised(pkg, oracle):
    """Two solves queued on one ctx on different streams share the ctx's work counter: the library orders them."""
    torch = _torch()
    prob = pkg.problems.quadrotor(20)
    B = 60_000
    dev = torch.device("cuda:0")
    xa, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    xb, _ = pkg.workloads.quadrotor_hover_batch(B, 2 * B, mult=0.25)
    ra = oracle.solve_batch(prob, xa, xref, dtype=np.float32, nthreads=8)
    rb = oracle.solve_batch(prob, xb, xref, dtype=np.float32, nthreads=8)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    xrd = torch.from_numpy(xref).to(dev)
    bufs = []
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    torch.cuda.synchronize()
    for x0, st_ in zip((xa, xb), streams):
        x0d = torch.from_numpy(x0).to(dev)
        u = torch.empty((B, 9, 4), device=dev)
        x = torch.empty((B, 10, 12), device=dev)
        it = torch.empty(B, dtype=torch.int32, device=dev)
        bufs.append((x0d, x, u, it))
    torch.cuda.synchronize()
    for (x0d, x, u, it), st_ in zip(bufs, streams):
        s.solve_raw(B, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, None, None, stream=st_.cuda_stream)
    torch.cuda.synchronize()
    for (x0d, x, u, it), r in zip(bufs, (ra, rb)):
        assert_same(it.cpu().numpy(), r.iter, "iter")
        assert_same(x.cpu().numpy(), r.x, "x")
        assert_same(u.cpu().numpy(), r.u, "u")


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_device_solve_equals_single_device(pkg, oracle, ndev):
    torch = _torch()
    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d devices" % ndev)
    prob = pkg.problems.quadrotor(20)
    B = 70_001                                  # odd: ranges of different length
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    ref = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=8)
    m = pkg.capi.Multi(prob, dtype=np.float32, policy="parity", devices=ndev)
    assert m.device_count == ndev
    out = m.solve(x0, xref, outputs=("x", "u", "u0", "iter", "status", "resid"))
    for k in ("iter", "status", "x", "u", "resid"):
        assert_same(out[k], getattr(ref, k), "multi %s" % k)
    assert_same(out["u0"], ref.u[:, 0, :], "multi u0")
    st = m.stats()
    assert st["instances"] == B and st["iterations"] == int(ref.iter.sum())
    used = [d for d in st["per_device"] if d["instances"]]
    assert len(used) == ndev
    # per-instance Xref + warm start through the multi path
    x0t, xrt = pkg.workloads.quadrotor_tracking_batch(0, 40_000)
    reft = oracle.solve_batch(prob, x0t, xrt, dtype=np.float32, want_state=True, nthreads=8)
    z = lambda *sh: np.zeros(sh, np.float32)
    warm = {"d": z(40_000, 9, 4), "y": z(40_000, 9, 4), "z": z(40_000, 9, 4), "g": z(40_000, 10, 12), "v": z(40_000, 10, 12)}
    outt = m.solve(x0t, xrt, warm=warm)
    for k in ("iter", "status", "x", "u"):
        assert_same(outt[k], getattr(reft, k), "multi tracking %s" % k)
    for k in ("d", "y", "z", "g", "v"):
        assert_same(outt["warm"][k], reft.state[k], "multi warm " + k)
    # a small batch stays on one device
    outs = m.solve(x0[:1000], xref)
    assert_same(outs["iter"], ref.iter[:1000], "small")
    assert sum(1 for d in m.stats()["per_device"] if d["instances"]) == 1
    m.close()


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_rollout_matches_single_device_batch(pkg, ndev):
    """tmpc_multi_rollout: the closed loop of one host batch over `ndev` devices (contiguous instance ranges, strided history
    slices) equals tmpc_batch_rollout of the whole batch on one device, element for element -- fixed Xref (fused loop) and a
    reference table with per-instance window starts."""
    import torch
    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d devices" % ndev)
    prob = pkg.problems.quadrotor(20)
    B, steps = 70_001, 5
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    table = np.ascontiguousarray(pkg.problems.quadrotor_trajectory().T).astype(np.float32)
    starts = (np.arange(B) % 295).astype(np.int32)
    m = pkg.capi.Multi(prob, dtype=np.float32, policy="parity", devices=ndev)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    for use_table in (False, True):
        hm = m.rollout(x0, steps, xref=None if use_table else xref, table=table if use_table else None,
                       start=starts if use_table else None, last=True)
        assert len([d for d in m.stats()["per_device"] if d["instances"]]) == ndev
        b = pkg.capi.Batch(s, B)
        b.set_x0(x0)
        if use_table:
            b.set_xref_table(table, starts)
        else:
            b.set_xref(xref)
        h1 = b.rollout(steps, reset_duals=True)
        for k in ("x0", "u0", "iter", "status"):
            assert_same(hm[k], h1[k], "table %s history %s" % (use_table, k))
        assert_same(hm["x"], b.get("x"), "last x")
        assert_same(hm["u"], b.get("u"), "last u")
        b.close()
    s.close()
