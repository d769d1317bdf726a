"""Device-resident cold solve of the hover workload with per-instance boxes (tmpc_set_instance_bounds): kernel time of the
run-time-shape kernel on that path.  usage: profile_instance_bounds.py [instances]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
capi = pkg.capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
prob = pkg.problems.quadrotor(20)
x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
rng = np.random.default_rng(0)
su = (0.5 * rng.uniform(0.6, 1.4, (B, 1, 1)) * np.ones((1, prob.N - 1, prob.nu))).astype(np.float32)
sx = (5.0 * rng.uniform(0.6, 1.4, (B, 1, 1)) * np.ones((1, prob.N, prob.nx))).astype(np.float32)
s = capi.Solver(prob, dtype=np.float32, policy="parity")
dev = torch.device("cuda:0")
x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
x = torch.empty((B, prob.N, prob.nx), dtype=torch.float32, device=dev)
u = torch.empty((B, prob.N - 1, prob.nu), dtype=torch.float32, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
rs = torch.empty((B, 4), dtype=torch.float32, device=dev)
for label in ("shared bounds (specialised kernel)", "per-instance bounds (run-time-shape kernel)"):
    if label.startswith("per"):
        s.set_instance_bounds(-sx, sx, -su, su)
    for _ in range(2):
        s.solve_raw(B, x0d, xrd, True, capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        q = s.stats()
    print("%s B=%d: %.3f ms  %.3e solves/s  %.3e it/s  mean iters %.2f" %
          (label, B, q["kernel_ms"], B / q["kernel_ms"] * 1e3, q["iterations"] / q["kernel_ms"] * 1e3, q["iterations"] / B), flush=True)
