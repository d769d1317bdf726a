"""A/B of library builds on the headline workload: python tools/ab_libs.py <lib.so> [<lib.so> ...]  (development helper).
Each library runs in its own process (TMPC_LIB_PATH); prints the best kernel time of 6 device-resident 1M-instance hover solves."""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from __graft_entry__ import load_package
pkg = load_package()
prob = pkg.problems.quadrotor(20)
B = 1 << 20
dev = torch.device("cuda:0")
for wl in sys.argv[1].split(","):
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25) if wl == "hover" else pkg.workloads.quadrotor_tracking_batch(0, B)
    s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
    x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
    x = torch.empty((B, 10, 12), device=dev); u = torch.empty((B, 9, 4), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); rs = torch.empty((B, 4), device=dev)
    ms = []
    for _ in range(6):
        s.solve_raw(B, x0d, xrd, xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ms.append(s.stats()["kernel_ms"])
    q = s.stats()
    print("   %-6s best %.3f ms  median %.3f ms  trips/iter %.4f  iters %d" % (wl, min(ms), sorted(ms)[3], q["trips"] / q["iterations"], q["iterations"]), flush=True)
    s.close()
'''
wl = os.environ.get("AB_WORKLOADS", "hover")
for lib in sys.argv[1:]:
    print(lib, flush=True)
    env = dict(os.environ, TMPC_LIB_PATH=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD, wl], env=env, check=False)
