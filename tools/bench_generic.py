"""Device-resident timing of the run-time-shape kernel (tmpc_kernel_rt.cuh) on a few shapes, plus the three BASELINE
shapes forced onto it (TMPC_KERNEL=rt) next to their specialised kernels.  One JSON line per case on stdout."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()


def flop_per_iter(n, m, N):
    S = m * (N - 1) + n * N
    mac = (N - 1) * (2 * n * n + 4 * n * m + m * m) + n * n + n * N + n
    elt = (N - 1) * (2 * m + 3 * n) + 11 * S + 2 * m * (N - 1) + 2 * n * N + n + 2
    return 2 * mac + elt


def run(name, prob, x0, xref, policy="parity", dtype=np.float32, reps=3):
    B = x0.shape[0]
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    s = pkg.capi.Solver(prob, dtype=dtype, policy=policy)
    dev = torch.device("cuda:0")
    x0d = torch.from_numpy(x0.astype(dtype)).to(dev)
    xrd = torch.from_numpy(xref.astype(dtype)).to(dev)
    x = torch.empty((B, prob.N, prob.nx), dtype=tdt, device=dev)
    u = torch.empty((B, prob.N - 1, prob.nu), dtype=tdt, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), dtype=tdt, device=dev)
    stream = torch.cuda.Stream()
    best = None
    for r in range(reps):
        s.solve_raw(B, x0d, xrd, xref.ndim == 2, pkg.capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        stt = s.stats()
        if best is None or stt["kernel_ms"] < best["kernel_ms"]:
            best = stt
    ms, iters = best["kernel_ms"], best["iterations"]
    f = flop_per_iter(prob.nx, prob.nu, prob.N)
    print(json.dumps({"name": name, "shape": [prob.nx, prob.nu, prob.N], "dtype": np.dtype(dtype).name, "policy": policy, "batch": B,
                      "kernel": os.environ.get("TMPC_KERNEL", "default"), "ms": round(ms, 3), "solves_per_s": B / ms * 1e3,
                      "iters_per_s": iters / ms * 1e3, "mean_iters": iters / B, "lanes": best["lanes"],
                      "tflops_algorithmic": iters * f / ms * 1e3 / 1e12}), flush=True)
    s.close()


def generic(nx, nu, N, B, **kw):
    prob = pkg.problems.random_system(nx, nu, N, seed=7 + nx)
    rng = np.random.default_rng(nx * 100 + nu)
    x0 = (rng.uniform(-2, 2, (B, nx))).astype(np.float32)
    x0[::2] *= np.float32(0.1)
    run("random_%d_%d_%d" % (nx, nu, N), prob, x0, np.zeros((N, nx), np.float32), **kw)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
    q = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    run("quadrotor_specialised", q, x0, xref)
    os.environ["TMPC_KERNEL"] = "rt"
    run("quadrotor_rt", q, x0, xref)
    run("quadrotor_rt", q, x0, xref, policy="fast")
    run("quadrotor_rt", q, x0, xref, dtype=np.float64)
    c = pkg.problems.cartpole()
    x0, xref = pkg.workloads.cartpole_batch(0, B * 4)
    run("cartpole_rt", c, x0, xref)
    os.environ.pop("TMPC_KERNEL")
    for shape in ((6, 3, 20), (10, 5, 10), (16, 8, 25), (24, 6, 10), (40, 10, 6), (64, 16, 8)):
        generic(*shape, B=max(B // 2, 4096))
    generic(16, 8, 25, B=max(B // 2, 4096), dtype=np.float64)
    # resident blocks per SM (scratch working set vs latency hiding)
    for per_sm in ("1", "2", "3"):
        os.environ["TMPC_RT_BLOCKS_PER_SM"] = per_sm
        os.environ["TMPC_KERNEL"] = "rt"
        x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
        run("quadrotor_rt_blocks_per_sm_" + per_sm, q, x0, xref)
        os.environ.pop("TMPC_KERNEL")
        generic(16, 8, 25, B=max(B // 2, 4096))
    os.environ.pop("TMPC_RT_BLOCKS_PER_SM")


if __name__ == "__main__":
    main()
