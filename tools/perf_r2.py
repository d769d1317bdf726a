"""Round-2 development probe: kernel times of the streamed-state paths of the fp32 12/4/10 kernel (device-resident, best of 3).
usage: perf_r2.py [hover,track,rollout,ib,warm]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
capi = pkg.capi
dev = torch.device("cuda:0")
prob = pkg.problems.quadrotor(20)
which = (sys.argv[1] if len(sys.argv) > 1 else "hover,track,rollout,ib,warm").split(",")
PEAK = 148 * 128 * 2 * 1.965e9


def bufs(B):
    return (torch.empty((B, 10, 12), device=dev), torch.empty((B, 9, 4), device=dev), torch.empty(B, dtype=torch.int32, device=dev),
            torch.empty(B, dtype=torch.int32, device=dev), torch.empty((B, 4), device=dev))


def solve(s, B, x0, xref, warm=None, reps=3, restore=None):
    x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
    x, u, it, st, rs = bufs(B)
    best = None
    for _ in range(reps):
        if restore:
            for k in warm:
                warm[k].copy_(restore[k])
        torch.cuda.synchronize()
        s.solve_raw(B, x0d, xrd, xref.ndim == 2, capi.TMPC_MEM_DEVICE, x, u, it, st, rs, warm=warm, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        q = s.stats()
        if best is None or q["kernel_ms"] < best["kernel_ms"]:
            best = q
    return best


def show(name, B, q):
    ms = q["kernel_ms"]
    print("%-34s B=%-8d %8.3f ms  %.3e solves/s  %.3e it/s  mean it %5.2f  trips/it %.3f  frac %.3f  sched %d" %
          (name, B, ms, B / ms * 1e3, q["iterations"] / ms * 1e3, q["iterations"] / B, q["trips"] / max(q["iterations"], 1),
           q["iterations"] * 11918 / (ms * 1e-3) / PEAK, q["scheduled"]), flush=True)


B = 1 << 20
if "hover" in which:
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    s = capi.Solver(prob, dtype=np.float32, policy="parity")
    show("hover parity", B, solve(s, B, x0, xref))
    s.close()
if "track" in which:
    x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B)
    for env in ({}, {"TMPC_NO_XR_SCRATCH": "1"}, {"TMPC_LPT": "1"}):
        os.environ.update(env)
        s = capi.Solver(prob, dtype=np.float32, policy="parity")
        show("tracking parity %r" % (env,), B, solve(s, B, x0, xref))
        s.close()
        for k in env:
            del os.environ[k]
if "warm" in which:
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    for env in ({}, {"TMPC_NO_WM_SCRATCH": "1"}):
        os.environ.update(env)
        s = capi.Solver(prob, dtype=np.float32, policy="parity")
        warm = {k: torch.zeros((B, 9, 4) if k in "dyz" else (B, 10, 12), device=dev) for k in ("d", "y", "g", "v", "z")}
        q1 = solve(s, B, x0, xref, warm=warm, reps=1)
        show("warm-capable cold solve %r" % (env,), B, q1)
        saved = {k: v.clone() for k, v in warm.items()}
        x1 = (x0 * np.float32(1.01)).astype(np.float32)
        show("warm re-solve (x0 * 1.01) %r" % (env,), B, solve(s, B, x1, xref, warm=warm, restore=saved))
        s.close()
        for k in env:
            del os.environ[k]
if "rollout" in which:
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    for env in ({}, {"TMPC_ROLL": "0"}, {"TMPC_TEST_MIRROR": "8"}):
        os.environ.update(env)
        s = capi.Solver(prob, dtype=np.float32, policy="parity")
        b = capi.Batch(s, B)
        steps, best = 10, None
        for _ in range(2):
            b.reset(); b.set_x0(x0); b.set_xref(xref)
            ith = torch.empty((steps, B), dtype=torch.int32, device=dev)
            s._check(s.lib.tmpc_batch_rollout(b._b, steps, 1, None, None, ith.data_ptr(), None, capi.TMPC_MEM_DEVICE), "rollout")
            ms = b.last_rollout_ms()
            if best is None or ms < best[0]:
                best = (ms, int(ith.sum().item()))
        ms, iters = best
        q = s.stats()
        print("rollout hover 10 steps %r: %.3f ms  %.3e MPC steps/s  %.3e it/s  mean it/step %.2f  frac %.3f  trips/it (last launch) %.3f" %
              (env, ms, B * steps / ms * 1e3, iters / ms * 1e3, iters / (B * steps), iters * 11918 / (ms * 1e-3) / PEAK,
               q["trips"] / max(1, q["iterations"])), flush=True)
        b.close(); s.close()
        for k in env:
            del os.environ[k]
if "ib" in which:
    Bi = 262144
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, Bi, mult=0.25)
    rng = np.random.default_rng(0)
    su = (0.5 * rng.uniform(0.6, 1.4, (Bi, 1, 1)) * np.ones((1, 9, 4))).astype(np.float32)
    sx = (5.0 * rng.uniform(0.6, 1.4, (Bi, 1, 1)) * np.ones((1, 10, 12))).astype(np.float32)
    s = capi.Solver(prob, dtype=np.float32, policy="parity")
    show("shared bounds", Bi, solve(s, Bi, x0, xref))
    s.set_instance_bounds(-sx, sx, -su, su)
    show("per-instance bounds, f32 kernel", Bi, solve(s, Bi, x0, xref))
    s.close()
    os.environ["TMPC_IB_RT"] = "1"
    s = capi.Solver(prob, dtype=np.float32, policy="parity")
    s.set_instance_bounds(-sx, sx, -su, su)
    show("per-instance bounds, rt kernel", Bi, solve(s, Bi, x0, xref, reps=1))
    s.close()
    del os.environ["TMPC_IB_RT"]
