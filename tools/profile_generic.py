"""One device-resident solve of a run-time-shape problem (profiling target for ncu): python tools/profile_generic.py nx nu N B"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
nx, nu, N, B = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (16, 8, 25, 131072)
prob = pkg.problems.random_system(nx, nu, N, seed=7 + nx)
rng = np.random.default_rng(nx * 100 + nu)
x0 = rng.uniform(-2, 2, (B, nx)).astype(np.float32)
x0[::2] *= np.float32(0.1)
xref = np.zeros((N, nx), np.float32)
s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
dev = torch.device("cuda:0")
x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
x = torch.empty((B, N, nx), device=dev); u = torch.empty((B, N - 1, nu), device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); rs = torch.empty((B, 4), device=dev)
for _ in range(2):
    s.solve_raw(B, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(s.stats())
