"""One fused closed-loop rollout of hover instances (profiling target for ncu): profile_roll.py [B] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prob = pkg.problems.quadrotor(20)
x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
b = pkg.capi.Batch(s, B)
b.set_x0(x0)
b.set_xref(xref)
ith = torch.empty((steps, B), dtype=torch.int32, device="cuda:0")
s._check(s.lib.tmpc_batch_rollout(b._b, steps, 1, None, None, ith.data_ptr(), None, pkg.capi.TMPC_MEM_DEVICE), "rollout")
q = s.stats()
print("rollout ms", b.last_rollout_ms(), "iterations", int(ith.sum().item()), q)
