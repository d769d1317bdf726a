"""One device-resident fp64 solve of the hover workload (profiling target for ncu).  usage: profile_f64.py [instances]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
prob = pkg.problems.quadrotor(20)
x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
s = pkg.capi.Solver(prob, dtype=np.float64, policy="parity")
dev = torch.device("cuda:0")
f = lambda a: torch.from_numpy(a.astype(np.float64)).to(dev)
x = torch.empty((B, 10, 12), dtype=torch.float64, device=dev); u = torch.empty((B, 9, 4), dtype=torch.float64, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); rs = torch.empty((B, 4), dtype=torch.float64, device=dev)
for _ in range(2):
    s.solve_raw(B, f(x0), f(xref), True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    q = s.stats()
    print(q, "it/s %.3e" % (q["iterations"] / q["kernel_ms"] * 1e3))
