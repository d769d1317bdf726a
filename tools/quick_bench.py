"""Quick device-resident timing of the persistent kernel on the BASELINE configs (development helper)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()


def run(policy, dtype, B, workload, mult=0.25, reps=3):
    if workload == "hover":
        prob = pkg.problems.quadrotor(20); x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=mult); flop = 11918
    elif workload == "track":
        prob = pkg.problems.quadrotor(20); x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B); flop = 11918
    else:
        prob = pkg.problems.cartpole(); x0, xref = pkg.workloads.cartpole_batch(0, B); flop = 1771
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    s = pkg.capi.Solver(prob, dtype=dtype, policy=policy)
    dev = torch.device("cuda:0")
    x0d = torch.from_numpy(x0.astype(dtype)).to(dev)
    xrd = torch.from_numpy(xref.astype(dtype)).to(dev)
    x = torch.empty((B, prob.N, prob.nx), dtype=tdt, device=dev)
    u = torch.empty((B, prob.N - 1, prob.nu), dtype=tdt, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), dtype=tdt, device=dev)
    stream = torch.cuda.Stream()
    best = None
    for r in range(reps):
        s.solve_raw(B, x0d, xrd, xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        stt = s.stats()
        if best is None or stt["kernel_ms"] < best["kernel_ms"]:
            best = stt
    ms, iters = best["kernel_ms"], best["iterations"]
    print("%-6s %-7s %-6s B=%-8d mult=%-4s: %8.3f ms  %.3e solves/s  %.3e it/s  mean it %6.2f  trips/iter %.3f  %6.2f TFLOP/s" %
          (policy, np.dtype(dtype).name, workload, B, mult, ms, B / ms * 1e3, iters / ms * 1e3, iters / B,
           best["trips"] / max(iters, 1), iters * flop / ms * 1e3 / 1e12), flush=True)
    s.close()


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    for mult in (0.1, 0.25, 1.0):
        run("parity", np.float32, B, "hover", mult)
        run("fast", np.float32, B, "hover", mult)
    if which == "all":
        run("parity", np.float32, B * 4, "hover", 0.25)
        run("parity", np.float32, B, "track")
        run("fast", np.float32, B, "track")
        run("parity", np.float64, B // 8, "hover", 0.25)
        run("parity", np.float32, B * 4, "cart")
        run("fast", np.float32, B * 4, "cart")
