"""Device-resident timing of every BASELINE.json config that is not the bench.py headline, plus the (f)-row paths
(closed loop on the device, per-instance systems).  One JSON line per measurement (kernel time = CUDA events around
the launch inside the library, best of 3).  Single GPU; the multi-GPU configs shard by index ranges exactly like
bench.py, so their per-GPU batch is what is timed here.

    python tools/bench_configs.py [scale]        scale divides every batch (default 1)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
capi = pkg.capi
DEV = torch.device("cuda:0")
PEAK_TF = 148 * 128 * 2 * 1.965e9 / 1e12
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
FLOP = {"q": 11918, "c": 1771, "l": 344058}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def outputs(B, prob, tdt=torch.float32):
    return (torch.empty((B, prob.N, prob.nx), dtype=tdt, device=DEV), torch.empty((B, prob.N - 1, prob.nu), dtype=tdt, device=DEV),
            torch.empty(B, dtype=torch.int32, device=DEV), torch.empty(B, dtype=torch.int32, device=DEV),
            torch.empty((B, 4), dtype=tdt, device=DEV))


def emit(name, config, B, st, shape, bytes_per_solve, extra=None):
    ms, iters = st["kernel_ms"], st["iterations"]
    tf = iters * FLOP[shape] / (ms * 1e-3) / 1e12
    gbs = B * bytes_per_solve / (ms * 1e-3) / 1e9
    rec = {"name": name, "config": config, "instances": B, "kernel_ms": ms, "solves_per_s": B / (ms * 1e-3), "iters_per_s": iters / (ms * 1e-3),
           "mean_iters": iters / B, "solved_frac": st["solved"] / B, "algorithmic_tflops": tf, "fp32_roofline_frac": tf / PEAK_TF,
           "algorithmic_gbs": gbs, "hbm_roofline_frac": gbs / HBM, "pattern": st.get("pattern", 0)}
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)


def timed_solve(s, B, x0d, xrd, shared, prob, warm_state=None, reps=3, restore=None):
    x, u, it, stt, rs = outputs(B, prob)
    stream = torch.cuda.Stream()
    best = None
    for _ in range(reps):
        if restore is not None:
            for k in warm_state:
                warm_state[k].copy_(restore[k])
        torch.cuda.synchronize()
        s.solve_raw(B, x0d, xrd, shared, capi.TMPC_MEM_DEVICE, x, u, it, stt, rs, warm=warm_state, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        q = s.stats()
        if best is None or q["kernel_ms"] < best["kernel_ms"]:
            best = q
    return best


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    quad = pkg.problems.quadrotor(20)
    # ---- config 3: tracking, per-instance reference windows (4M over 2/4/8 GPUs -> per-GPU shard of 1M timed here)
    B = (1 << 20) // scale
    x0, xref = pkg.workloads.quadrotor_tracking_batch(0, B)
    for pol in ("parity", "fast"):
        s = capi.Solver(quad, dtype=np.float32, policy=pol)
        emit("config3_tracking_" + pol, "quadrotor_tracking, per-instance Xref windows, 1/4 of the 4M batch (one shard of 4 GPUs)", B,
             timed_solve(s, B, dev(x0), dev(xref), False, quad), "q", 1160)
        s.close()
    # ---- config 4: cartpole 16M over 8 GPUs -> 2M per GPU
    cart = pkg.problems.cartpole()
    B = (1 << 21) // scale
    x0, xref = pkg.workloads.cartpole_batch(0, B)
    for pol in ("parity", "fast"):
        s = capi.Solver(cart, dtype=np.float32, policy=pol)
        emit("config4_cartpole_" + pol, "codegen_cartpole 4/1/10, 1/8 of the 16M batch (one shard of 8 GPUs)", B,
             timed_solve(s, B, dev(x0), dev(xref), True, cart), "c", 220)
        s.close()
    # ---- config 5: 32/8/50, cold solve then x0 perturbed by 1 % and re-solved with {d,y,g,v,z} in HBM
    big = pkg.problems.random_system()
    B = 32768 // scale
    x0, xref = pkg.workloads.random_system_batch(0, B)
    for pol in ("parity", "fast"):
        s = capi.Solver(big, dtype=np.float32, policy=pol)
        warm = {k: torch.zeros((B, 49, 8) if k in "dyz" else (B, 50, 32), device=DEV) for k in ("d", "y", "g", "v", "z")}
        cold = timed_solve(s, B, dev(x0), dev(xref), True, big, warm_state=warm, reps=1)
        emit("config5_large_cold_" + pol, "random 32/8/50 system, cold solve writing the warm state (untimed leg of config 5)", B, cold, "l",
             128 + 6400 + 1568 + 8 + 17504)
        saved = {k: v.clone() for k, v in warm.items()}
        x1 = pkg.workloads.perturb_x0(x0, 0)
        emit("config5_large_warm_" + pol, "random 32/8/50 system, x0 perturbed 1 %, warm start read from / written to HBM (the timed leg)", B,
             timed_solve(s, B, dev(x1), dev(xref), True, big, warm_state=warm, restore=saved), "l", 128 + 6400 + 1568 + 8 + 2 * 17504)
        s.close()
    # ---- (f)2: the examples' closed loop on the device: 10 MPC steps of 1M hover instances, duals reset each step
    B = (1 << 20) // scale
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    s = capi.Solver(quad, dtype=np.float32, policy="parity")
    b = capi.Batch(s, B)
    steps = 10
    best = None
    for _ in range(2):
        b.reset(); b.set_x0(x0); b.set_xref(xref)
        it_h = torch.empty((steps, B), dtype=torch.int32, device=DEV)
        s._check(s.lib.tmpc_batch_rollout(b._b, steps, 1, None, None, it_h.data_ptr(), None, capi.TMPC_MEM_DEVICE), "rollout")
        ms = b.last_rollout_ms()
        if best is None or ms < best[0]:
            best = (ms, int(it_h.sum().item()))
    ms, iters = best
    print(json.dumps({"name": "closed_loop_rollout_hover", "config": "10 closed-loop MPC steps (window, reset duals, warm solve, plant) on the device",
                      "instances": B, "steps": steps, "rollout_ms": ms, "mpc_steps_per_s": B * steps / (ms * 1e-3), "iters_per_s": iters / (ms * 1e-3),
                      "mean_iters_per_step": iters / (B * steps)}), flush=True)
    b.close(); s.close()
    # ---- (f)1: per-instance systems: batched precompute + solve with per-instance models
    import time
    B = (1 << 18) // scale
    rng = np.random.default_rng(1)
    A = np.repeat(quad.Adyn[None], B, 0).copy()
    off = ~np.eye(12, dtype=bool)
    A[:, off] *= (1.0 + 0.2 * rng.uniform(-1, 1, (B, 1)))
    Bm = quad.Bdyn[None] * (1.0 + 0.3 * rng.uniform(-1, 1, (B, 1, 1)))
    Q = np.repeat(quad.Q[None], B, 0); R = np.repeat(quad.R[None], B, 0); rho = 5.0 * (1.0 + 0.4 * rng.uniform(-1, 1, B))
    s = capi.Solver(quad, dtype=np.float32, policy="parity")
    t0 = time.perf_counter()
    sy = capi.Systems(s, A, Bm, Q, R, rho)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sw = sy.get("sweeps")
    print(json.dumps({"name": "systems_precompute", "config": "batched Riccati precompute (double), perturbed quadrotor models, incl. H2D of the models",
                      "instances": B, "seconds": t1 - t0, "systems_per_s": B / (t1 - t0), "mean_sweeps": float(sw.mean()),
                      "riccati_sweeps_per_s": float(sw.sum()) / (t1 - t0)}), flush=True)
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=0.25)
    x, u, it, stt, rs = outputs(B, quad)
    x0d, xrd = dev(x0), dev(xref)
    best = None
    for _ in range(3):
        sy.solve_raw(x0d, xrd, True, x, u, it, stt, rs)
        torch.cuda.synchronize()
        q = s.stats()
        if best is None or q["kernel_ms"] < best["kernel_ms"]:
            best = q
    emit("systems_solve_parity", "per-instance models: each lane's 496 loop coefficients resident in tensor memory (TMPC_KERNEL=sys_global: re-read "
         "from its global block every stage)", B, best, "q", 680 + 4 * 960, extra={"kernel_variant": os.environ.get("TMPC_KERNEL", "sys_tmem")})


if __name__ == "__main__":
    main()
