"""One device-resident solve of a BASELINE config (profiling target for ncu).
usage: profile_cfg.py <hover|track|cartpole> [policy] [instances]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
which = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
policy = sys.argv[2] if len(sys.argv) > 2 else "parity"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
if which == "hover":
    prob = pkg.problems.quadrotor(20); x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
elif which == "track":
    prob = pkg.problems.quadrotor(20); x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B)
else:
    prob = pkg.problems.cartpole(); x0, xref = pkg.workloads.cartpole_batch(0, B)
s = pkg.capi.Solver(prob, dtype=np.float32, policy=policy)
dev = torch.device("cuda:0")
x0d = torch.from_numpy(x0).to(dev)
xrd = torch.from_numpy(xref).to(dev)
x = torch.empty((B, prob.N, prob.nx), dtype=torch.float32, device=dev)
u = torch.empty((B, prob.N - 1, prob.nu), dtype=torch.float32, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
rs = torch.empty((B, 4), dtype=torch.float32, device=dev)
for _ in range(3):
    s.solve_raw(B, x0d, xrd, xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    stt = s.stats()
    print("%s %s B=%d: %.3f ms  %.3e solves/s  %.3e it/s  trips/iter %.3f" %
          (which, policy, B, stt["kernel_ms"], B / stt["kernel_ms"] * 1e3, stt["iterations"] / stt["kernel_ms"] * 1e3,
           stt["trips"] / max(stt["iterations"], 1)), flush=True)
