"""Cartpole closed loop (config 4 in the regime BASELINE names), fused vs one launch per step: perf_cart_roll.py [B] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prob = pkg.problems.cartpole()
x0, xref = pkg.workloads.cartpole_batch(0, B)
x0 = (0.2 * x0).astype(np.float32)
for env in ({}, {"TMPC_ROLL": "0"}, {"TMPC_ROLL_ASYNC_REFILL": "1"}, {"TMPC_KERNEL": "small384"}):
    os.environ.update(env)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    b = pkg.capi.Batch(s, B)
    b.set_x0(x0); b.set_xref(xref)
    ith = torch.empty((steps, B), dtype=torch.int32, device="cuda:0")
    s._check(s.lib.tmpc_batch_rollout(b._b, 4, 1, None, None, None, None, pkg.capi.TMPC_MEM_DEVICE), "rollout")
    best = None
    for _ in range(2):
        s._check(s.lib.tmpc_batch_rollout(b._b, steps, 1, None, None, ith.data_ptr(), None, pkg.capi.TMPC_MEM_DEVICE), "rollout")
        ms = b.last_rollout_ms()
        best = ms if best is None else min(best, ms)
    iters = int(ith.sum().item())
    q = s.stats()
    print("cartpole closed loop %r: %.3f ms per %d steps  %.3e MPC steps/s  %.3e it/s  mean it/step %.2f  fp32 frac %.3f  trips/it %.3f" %
          (env, best, steps, B * steps / best * 1e3, iters / best * 1e3, iters / (B * steps), iters * 1771 / (best * 1e-3) / (148 * 128 * 2 * 1.965e9),
           q["trips"] / max(1, q["iterations"])), flush=True)
    b.close(); s.close()
    for k in env:
        del os.environ[k]
