"""One device-resident cold solve of the 32/8/50 system (profiling target for ncu).  usage: profile_large.py [instances]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
prob = pkg.problems.random_system()
x0, xref = pkg.workloads.random_system_batch(0, B)
s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
dev = torch.device("cuda:0")
x = torch.empty((B, 50, 32), device=dev); u = torch.empty((B, 49, 8), device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); rs = torch.empty((B, 4), device=dev)
s.solve_raw(B, torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev), True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs,
            stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(s.stats())
