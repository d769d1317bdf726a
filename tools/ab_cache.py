"""A/B of where the shared model/cache lives in the fp32 12/4/10 kernel: constant bank (default) vs shared memory staged by one TMA
bulk copy per CTA (TMPC_KERNEL=f32_tma_cache).  Headline workload, quadrotor pattern and dense instance; prints kernel ms (best of 5)."""
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
x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
s = pkg.capi.Solver(prob, dtype=np.float32, policy="parity")
x0d, xrd = torch.from_numpy(x0).to(dev), torch.from_numpy(xref).to(dev)
x = torch.empty((B, 10, 12), device=dev); u = torch.empty((B, 9, 4), device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev); rs = torch.empty((B, 4), device=dev)
ms = []
for _ in range(5):
    s.solve_raw(B, x0d, xrd, True, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ms.append(s.stats()["kernel_ms"])
q = s.stats()
print("%-34s pattern %d: best %.3f ms  iters %d  checksum %d %.6f" % (sys.argv[1], q["pattern"], min(ms), q["iterations"], int(it.sum()), float(u.double().sum())), flush=True)
'''
for dense in ("", "1"):
    for kern in ("", "f32_tma_cache"):
        env = dict(os.environ)
        if dense:
            env["TMPC_DENSE"] = "1"
        if kern:
            env["TMPC_KERNEL"] = kern
        subprocess.run([sys.executable, "-c", CHILD, "%s %s" % (kern or "constant bank (default)", "dense" if dense else "quadrotor pattern")], env=env, check=False)
