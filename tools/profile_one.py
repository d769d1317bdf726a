"""One device-resident solve of config 2 (profiling target for ncu)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
policy = sys.argv[1] if len(sys.argv) > 1 else "parity"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
mult = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
prob = pkg.problems.quadrotor(20)
x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=mult)
s = pkg.capi.Solver(prob, dtype=np.float32, policy=policy)
dev = torch.device("cuda:0")
x0d = torch.from_numpy(x0).to(dev)
xrd = torch.from_numpy(xref).to(dev)
x = torch.empty((B, 10, 12), dtype=torch.float32, device=dev)
u = torch.empty((B, 9, 4), dtype=torch.float32, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
rs = torch.empty((B, 4), dtype=torch.float32, device=dev)
for _ in range(2):
    s.solve_raw(B, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(s.stats())
