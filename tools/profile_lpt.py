"""Launch list helper: one device-resident hover solve and one tracking solve (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
prob = pkg.problems.quadrotor(20)
B = 1 << 20
dev = torch.device("cuda:0")
for wl in ("hover", "track"):
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25) if wl == "hover" else pkg.workloads.quadrotor_tracking_batch(0, B)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
    x = torch.empty((B, 10, 12), device=dev); u = torch.empty((B, 9, 4), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); rs = torch.empty((B, 4), device=dev)
    for _ in range(2):
        s.solve_raw(B, x0d, xrd, xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    print(wl, s.stats())
