"""Per-instance box bounds (tmpc_set_instance_bounds: the wrapper's set_xmin / set_xmax / set_umin / set_umax,
tiny_wrapper.cpp:43-129, with a leading batch dimension) through the C ABI, bit-exact (PARITY policy) against the CPU
oracle solving every instance on its own with that instance's bounds in the problem (admm.cpp:53,59)."""
import copy

import numpy as np
import pytest

from conftest import assert_same

pytestmark = pytest.mark.gpu


def _boxes(prob, B, rng, dtype):
    """Boxes that differ per instance, per stage and per coordinate; tight enough that many projections are active."""
    nx, nu, N = prob.nx, prob.nu, prob.N
    bu = 1.0 if prob.u_max is None else np.abs(np.asarray(prob.u_max, dtype=np.float64))[None]
    bx = 1.0 if prob.x_max is None else np.abs(np.asarray(prob.x_max, dtype=np.float64))[None]
    su = bu * rng.uniform(0.4, 1.6, (B, 1, 1)) * rng.uniform(0.8, 1.2, (B, N - 1, nu))
    sx = bx * rng.uniform(0.3, 1.2, (B, 1, 1)) * rng.uniform(0.8, 1.2, (B, N, nx))
    return (-sx).astype(dtype), (sx * 0.9).astype(dtype), (-su).astype(dtype), (su * 1.1).astype(dtype)


def _oracle_each(oracle, prob, x0, xref, b, dtype, warm=None, want_state=False, **settings):
    xmin, xmax, umin, umax = b
    outs = []
    for i in range(x0.shape[0]):
        p = copy.copy(prob)
        p.x_min, p.x_max, p.u_min, p.u_max = xmin[i], xmax[i], umin[i], umax[i]
        for k, v in settings.items():
            setattr(p, k, v)
        w = None if warm is None else {k: warm[k][i:i + 1] for k in warm}
        xr = xref if xref.ndim == 2 else xref[i]
        outs.append(oracle.solve_batch(p, x0[i:i + 1], xr, dtype=dtype, warm=w, want_state=want_state))
    cat = lambda name: np.concatenate([getattr(o, name) for o in outs])
    res = {k: cat(k) for k in ("iter", "status", "x", "u", "resid")}
    if want_state or warm is not None:
        res["state"] = {k: np.concatenate([o.state[k] for o in outs]) for k in ("d", "y", "g", "v", "z")}
    return res


def _cmp(out, ref, what):
    for name in ("iter", "status", "x", "u", "resid"):
        assert_same(out[name], ref[name], what + " " + name)


@pytest.mark.parametrize("which,dtype,B", [("quadrotor", np.float32, 161), ("quadrotor", np.float64, 67), ("cartpole", np.float32, 130),
                                           ("random_6_3_20", np.float32, 97), ("random_9_4_7", np.float64, 65)])
def test_instance_bounds_cold_host(pkg, oracle, which, dtype, B):
    if which == "quadrotor":
        prob = pkg.problems.quadrotor(20)
        x0, xref = pkg.workloads.quadrotor_hover_batch(3, 3 + B, mult=0.5)
    elif which == "cartpole":
        prob = pkg.problems.cartpole()
        x0, xref = pkg.workloads.cartpole_batch(0, B)
    else:
        nx, nu, N = (int(t) for t in which.split("_")[1:])
        prob = pkg.problems.random_system(nx, nu, N, seed=5)
        r = np.random.default_rng(1)
        x0 = r.uniform(-2, 2, (B, nx)).astype(np.float32)
        xref = r.uniform(-0.3, 0.3, (B, N, nx)).astype(np.float32)
    rng = np.random.default_rng(11)
    b = _boxes(prob, B, rng, dtype)
    s = pkg.capi.Solver(prob, dtype=dtype, policy="parity")
    shared = s.solve(x0, xref)
    s.set_instance_bounds(*b)
    out = s.solve(x0, xref)
    ref = _oracle_each(oracle, prob, x0, xref, b, dtype)
    _cmp(out, ref, which)
    assert len(set(ref["iter"].tolist())) > 1
    assert not np.array_equal(out["u"], shared["u"])            # the boxes really changed the answers
    st = s.stats()
    assert st["instances"] == B and st["iterations"] == int(ref["iter"].sum())
    # a batch of another size is refused while the bounds are set
    with pytest.raises(pkg.capi.TmpcError, match="batch differs"):
        s.solve(x0[:B - 1], xref if xref.ndim == 2 else xref[:B - 1])
    # state bounds switched off: only the per-instance input boxes act
    prob2 = copy.copy(prob)
    prob2.en_state_bound = 0
    s.set_model(prob2)
    out2 = s.solve(x0, xref)
    _cmp(out2, _oracle_each(oracle, prob, x0, xref, b, dtype, en_state_bound=0), which + " en_state_bound=0")
    # cleared: back to the shared bounds, bit for bit
    s.set_model(prob)
    s.set_instance_bounds()
    again = s.solve(x0, xref)
    for k in ("iter", "status", "x", "u", "resid"):
        assert_same(again[k], shared[k], which + " cleared " + k)
    s.close()


def test_instance_bounds_device_warm(pkg, oracle):
    """Device-resident buffers, warm start in place: second solve from the state the first one left."""
    import torch
    capi = pkg.capi
    B, dt = 96, np.float32
    prob = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(5, 5 + B, mult=0.5)
    b = _boxes(prob, B, np.random.default_rng(2), dt)
    s = capi.Solver(prob, dtype=dt, policy="parity")
    s.set_instance_bounds(*b)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nx, nu, N = prob.nx, prob.nu, prob.N
    warm = {k: torch.zeros((B, N - 1, nu) if k in "dyz" else (B, N, nx), dtype=torch.float32, device=dev) for k in ("d", "y", "g", "v", "z")}
    x, u = torch.empty((B, N, nx), dtype=torch.float32, device=dev), torch.empty((B, N - 1, nu), dtype=torch.float32, device=dev)
    it, stt = torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), dtype=torch.float32, device=dev)
    xrd = t(xref)
    state = None
    for rnd, scale in enumerate((1.0, 1.02)):
        x0r = (x0 * np.float32(scale)).astype(np.float32)
        s.solve_raw(B, t(x0r), xrd, True, capi.TMPC_MEM_DEVICE, x, u, it, stt, rs, warm=warm,
                    stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = _oracle_each(oracle, prob, x0r, xref, b, dt, warm=state, want_state=True)
        out = {"x": x.cpu().numpy(), "u": u.cpu().numpy(), "iter": it.cpu().numpy(), "status": stt.cpu().numpy(), "resid": rs.cpu().numpy()}
        _cmp(out, ref, "round %d" % rnd)
        for k in warm:
            assert_same(warm[k].cpu().numpy(), ref["state"][k], "round %d state %s" % (rnd, k))
        state = ref["state"]
    s.close()


def test_instance_bounds_wrapper_calls(pkg, oracle):
    """The wrapper's sequence with a batch dimension (tiny_wrapper.cpp:5-176): set_x0 / set_xref / set_*min,max per instance /
    call_tiny_solve / get_x / get_u, then reset_dual_variables and a second solve from the workspace the first one left."""
    capi = pkg.capi
    B, dt = 131, np.float32
    prob = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(9, 9 + B, mult=0.5)
    boxes = _boxes(prob, B, np.random.default_rng(4), dt)
    s = capi.Solver(prob, dtype=dt, policy="parity")
    s.set_instance_bounds(*boxes)
    b = capi.Batch(s, B)
    b.set_x0(x0)
    b.set_xref(xref)
    b.solve()
    r1 = _oracle_each(oracle, prob, x0, xref, boxes, dt, want_state=True)
    _cmp({k: b.get(k) for k in ("iter", "status", "x", "u", "resid")}, r1, "first solve")
    for k in ("d", "y", "g", "v", "z"):
        assert_same(b.get(k), r1["state"][k], "workspace." + k)
    x1 = (x0 * np.float32(0.97)).astype(np.float32)
    b.set_x0(x1)
    b.reset_dual_variables()
    b.solve()
    warm = {k: r1["state"][k].copy() for k in ("d", "y", "g", "v", "z")}
    warm["y"][:] = 0
    warm["g"][:] = 0
    r2 = _oracle_each(oracle, prob, x1, xref, boxes, dt, warm=warm, want_state=True)
    _cmp({k: b.get(k) for k in ("iter", "status", "x", "u", "resid")}, r2, "second solve")
    b.close()
    s.close()


def test_instance_bounds_large_batch_on_the_specialised_kernel(pkg, oracle):
    """fp32 12/4/10: per-instance boxes run on the specialised kernel (each lane copies its instance's box into its coalesced
    scratch rows at refill).  A batch large enough for several refills per lane and for the longest-expected-first schedule
    must agree bit for bit with the run-time-shape kernel on the same boxes (TMPC_IB_RT=1), and a prefix with the oracle."""
    import os
    prob = pkg.problems.quadrotor(20)
    B, dt = 90_000, np.float32
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    b = _boxes(prob, B, np.random.default_rng(5), dt)
    s = pkg.capi.Solver(prob, dtype=dt, policy="parity")
    s.set_instance_bounds(*b)
    out = s.solve(x0, xref)
    s.close()
    os.environ["TMPC_IB_RT"] = "1"
    try:
        s2 = pkg.capi.Solver(prob, dtype=dt, policy="parity")
        s2.set_instance_bounds(*b)
        out_rt = s2.solve(x0, xref)
        s2.close()
    finally:
        del os.environ["TMPC_IB_RT"]
    for k in ("iter", "status", "x", "u", "resid"):
        assert_same(out[k], out_rt[k], "f32 kernel vs run-time-shape kernel " + k)
    n = 1500
    ref = _oracle_each(oracle, prob, x0[:n], xref, tuple(a[:n] for a in b), dt)
    _cmp({k: out[k][:n] for k in ("iter", "status", "x", "u", "resid")}, ref, "prefix")
