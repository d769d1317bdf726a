"""One device-resident solve with per-instance systems (profiling target for ncu): batched precompute of perturbed
quadrotor models, then tmpc_solve_systems on the same hover workload as tools/bench_configs.py's systems line.
usage: profile_systems.py [instances] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
capi = pkg.capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
quad = pkg.problems.quadrotor(20)
rng = np.random.default_rng(1)
A = np.repeat(quad.Adyn[None], B, 0).copy()
off = ~np.eye(12, dtype=bool)
A[:, off] *= (1.0 + 0.2 * rng.uniform(-1, 1, (B, 1)))
Bm = quad.Bdyn[None] * (1.0 + 0.3 * rng.uniform(-1, 1, (B, 1, 1)))
Q = np.repeat(quad.Q[None], B, 0)
R = np.repeat(quad.R[None], B, 0)
rho = 5.0 * (1.0 + 0.4 * rng.uniform(-1, 1, B))
s = capi.Solver(quad, dtype=np.float32, policy="parity")
sy = capi.Systems(s, A, Bm, Q, R, rho)
dev = torch.device("cuda:0")
x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
x0d = torch.from_numpy(x0).to(dev)
xrd = torch.from_numpy(xref).to(dev)
x = torch.empty((B, quad.N, quad.nx), dtype=torch.float32, device=dev)
u = torch.empty((B, quad.N - 1, quad.nu), dtype=torch.float32, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
rs = torch.empty((B, 4), dtype=torch.float32, device=dev)
for _ in range(reps):
    sy.solve_raw(x0d, xrd, True, x, u, it, st, rs)
    torch.cuda.synchronize()
    q = s.stats()
    print("systems B=%d: %.3f ms  %.3e solves/s  %.3e it/s  trips/iter %.3f" %
          (B, q["kernel_ms"], B / q["kernel_ms"] * 1e3, q["iterations"] / q["kernel_ms"] * 1e3,
           q["trips"] / max(q["iterations"], 1)), flush=True)
