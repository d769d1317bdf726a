import os, sys, json
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo")
import numpy as np
import bench_generic as bg
B = 1 << 19
q = bg.pkg.problems.quadrotor(20)
x0, xref = bg.pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
os.environ["TMPC_KERNEL"] = "rt"
bg.run("quadrotor_rt", q, x0, xref)
bg.run("quadrotor_rt", q, x0, xref, dtype=np.float64)
os.environ.pop("TMPC_KERNEL")
bg.generic(16, 8, 25, B=B // 2)
bg.generic(40, 10, 6, B=B // 2)
bg.generic(6, 3, 20, B=B // 2)
