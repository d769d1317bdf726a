import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
pkg = load_package(); capi = pkg.capi
prob = pkg.problems.quadrotor(20); B = 1 << 20
track = len(sys.argv) > 1 and sys.argv[1] == "track"
x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B) if track else pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
s = capi.Solver(prob, dtype=np.float32, policy="parity")
hx0 = torch.from_numpy(x0).pin_memory(); hxr = torch.from_numpy(xref).pin_memory()
hx = torch.empty((B, 10, 12)).pin_memory(); hu = torch.empty((B, 9, 4)).pin_memory()
hit = torch.empty(B, dtype=torch.int32).pin_memory(); hst = torch.empty(B, dtype=torch.int32).pin_memory(); hrs = torch.empty((B, 4)).pin_memory()
for _ in range(2): s.solve_raw(B, hx0, hxr, not track, capi.TMPC_MEM_HOST, hx, hu, hit, hst, hrs)
t = time.perf_counter()
n = 8
for _ in range(n): s.solve_raw(B, hx0, hxr, not track, capi.TMPC_MEM_HOST, hx, hu, hit, hst, hrs)
dt = (time.perf_counter() - t) / n
print("track" if track else "hover", "overlap" if not os.environ.get("TMPC_NO_H2D_OVERLAP") else "no-overlap", "shift", os.environ.get("TMPC_D2H_SHIFT", "16"), "e2e ms %.3f  solves/s %.4e  kernel_ms %.3f" % (dt * 1e3, B / dt, s.stats()["kernel_ms"]), flush=True)
