#!/usr/bin/env python3
"""Benchmark of the batched TinyMPC ADMM hot path (BASELINE.json metric: batched MPC solves/sec, quadrotor
nx12 nu4 N10; % of the FP32 roofline).

  python bench.py --gpus N --steps K --warmup W            our arm  (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            the reference's own CPU tiny_solve on the host cores

A "step" = one tmpc_solve over one batch: configs[1] of BASELINE.json -- 1,048,576 quadrotor_hovering instances
with seeded random initial states (mult 0.25), one shared cache and Xref, cold start, per GPU (weak
scaling: rank r owns instance indices [r*B, (r+1)*B)).  `value` is whole-job solves/s with x0 resident in HBM
and x/u/iter/status/resid written to HBM; `e2e` is the same through the C ABI with pinned HOST buffers
(H2D of x0 and D2H of every output inside the timed region).  PyTorch is used for device buffers, events,
torch.distributed and nothing else.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

FLOP_PER_ITER = {"q": 11918, "c": 1771}     # 2*MAC + ELT
BYTES_PER_SOLVE = {"q": 680, "c": 220}      # algorithmic HBM bytes per solve, fp32, shared Xref


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="instances per GPU per step")
    ap.add_argument("--mult", type=float, default=0.25, help="initial-state spread (SURVEY 8d)")
    ap.add_argument("--policy", default="parity", choices=["parity", "fast"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = int(float(f[1]))
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    sm.append(int(float(f[0])))
                    for n, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                continue
        if not sm and self.rows:
            try:
                sm = [int(float(self.rows[-1][1].split(",")[0]))]
            except ValueError:
                pass
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(pkg, args, prob, quiet=False):
    """Time the reference's own tiny_solve (oracle/_ref, compiled from /root/reference) looping over a bounded
    prefix of the same workload on all host cores.  Falls back to the plain-C port when _ref is absent."""
    from oracle.pyoracle import OracleLib, RefLib
    cores = os.cpu_count() or 1
    kind, lib, build = None, None, None
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    for cfg, need, desc in (("q_f32_v3", ("avx2", "fma"), "-O3 -mavx2 -mfma"), ("q_f32", (), "-O3 (SSE2)")):
        if RefLib.available(cfg) and all(n in flags for n in need):
            kind, lib, build = "reference", RefLib(cfg), "oracle/_ref/libref_%s.so: reference admm.cpp + vendored Eigen, g++ %s" % (cfg, desc)
            break
    if lib is None:
        kind, lib, build = "port", OracleLib(), "oracle/libtinympc_oracle.so (plain-C restatement, gcc -O2)"
    W = pkg.workloads

    def run(n):
        x0, xref = W.quadrotor_hover_batch(0, n, mult=args.mult)
        t = time.perf_counter()
        if kind == "reference":
            r = lib.solve_batch(prob, x0, xref, nthreads=cores)
        else:
            r = lib.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=cores)
        return time.perf_counter() - t, r

    n = 4000 * cores
    dt, r = run(n)                                   # calibration (also warms the threads)
    n = int(max(n, min(n * args.cpu_seconds / max(dt, 1e-3), 4_000_000)))
    dt, r = run(n)
    return {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": kind,
            "sample": "first %d instances of the same seeded workload (mult %.2f), %.1f s, %s, float" % (n, args.mult, dt, build),
            "iters_per_s": float(r.iter.sum()) / dt, "mean_iters": float(r.iter.mean()), "seconds": dt}, n, dt


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything any library prints to fd 1 during the run
    (torch.distributed's NCCL banner, for one) has been routed to stderr by main()."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = load_package()
    prob = pkg.problems.quadrotor(20)
    config = {"workload": "batched quadrotor_hovering (BASELINE configs[1]): nx=12 nu=4 N=10, 20 Hz cache, "
                          "box bounds |u|<=0.5 |x|<=5, rho=5, tol 1e-3, max_iter 100, shared cache+Xref, cold start",
              "instances_per_gpu": args.batch, "x0_spread_mult": args.mult, "seed": 1234, "policy": args.policy,
              "sharding": "index ranges, no data-path collective",
              "l2": "each step writes %.0f MB of outputs per GPU (> 126 MB L2), so the %.0f MB of x0 cannot stay "
                    "L2-resident between steps" % (args.batch * 680 / 1e6, args.batch * 48 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        base, n, dt = cpu_reference(pkg, args, prob)
        steps = max(1, args.steps)
        # K bounded-sample steps after W warm-up steps (the calibration above is warm-up)
        vals = [base["value"]]
        for _ in range(min(steps - 1, 2)):
            b2, _, _ = cpu_reference(pkg, argparse.Namespace(**{**vars(args), "cpu_seconds": min(args.cpu_seconds, 6.0)}), prob)
            vals.append(b2["value"])
        v = float(np.mean(vals))
        out = {"impl": "reference", "metric": "batched MPC solves/sec (quadrotor nx12 nu4 N10)", "value": v,
               "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config,
               "cpu_baseline": {**base, "value": v},
               "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(json.dumps(out))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi = pkg.capi
    B = args.batch
    b0, b1 = pkg.sharding.shard_range(rank, world, per_rank=B)
    x0_np, xref_np = pkg.workloads.quadrotor_hover_batch(b0, b1, mult=args.mult)
    solver = capi.Solver(prob, dtype=np.float32, policy=args.policy, device=local)
    x0 = torch.from_numpy(x0_np).to(dev)
    xref = torch.from_numpy(xref_np).to(dev)
    x = torch.empty((B, prob.N, prob.nx), dtype=torch.float32, device=dev)
    u = torch.empty((B, prob.N - 1, prob.nu), dtype=torch.float32, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    rs = torch.empty((B, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)   # a real (non-NULL) stream: NULL would mean "the ctx's own stream"
    torch.cuda.set_stream(stream)

    def step():
        solver.solve_raw(B, x0, xref, True, capi.TMPC_MEM_DEVICE, x, u, it, st, rs, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    kernel_ms, launches = 0.0, 0
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    stats = solver.stats()           # statistics + CUDA-event time of the LAST step's kernel launch
    kernel_ms = stats["kernel_ms"]
    launches = stats["launches"] * args.steps

    # whole-job numbers: max time over ranks, totals over ranks (NCCL only for this statistics gather)
    vec = pkg.sharding.local_stats(it.cpu().numpy(), st.cpu().numpy(), prob.max_iter)
    assert vec[0] == stats["iterations"] and vec[1] == stats["solved"], "kernel statistics disagree with the outputs"
    vec, tmax = pkg.sharding.gather_stats(vec, [ms, kernel_ms], dist if world > 1 else None, dev)
    ms, kernel_ms = float(tmax[0]), float(tmax[1])
    total_iters, total_solved, total_inst = float(vec[0]), float(vec[1]), float(vec[2])
    hist = vec[3:]
    value = total_inst * args.steps / (ms * 1e-3)

    # ---- roofline of the one kernel (per GPU, per launch; kernel time by CUDA events inside the library,
    #      recorded on the launching stream around the launch)
    prop = torch.cuda.get_device_properties(local)
    sm_mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965
    peak_fp32 = prop.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12          # TFLOP/s at the clock seen under load
    peak_fp32_max = prop.multi_processor_count * 128 * 2 * (clocks["sm_max_mhz"] or 1965) * 1e6 / 1e12
    flops = stats["iterations"] * FLOP_PER_ITER["q"]
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = stats["instances"] * BYTES_PER_SOLVE["q"] / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes (read + write) of this kernel at this workload, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        if tj.get("batch") == args.batch and tj.get("policy") == args.policy and abs(tj.get("mult", -1) - args.mult) < 1e-9:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    # With the quadrotor structure specialisation (stats.pattern == 1) the kernel drops the terms whose coefficient is an
    # exact zero (118 of Adyn's 144, 40 of AmBKt's 144) and the multiplies by exact ones (Adyn's diagonal): per stage
    # 28 instead of 276 FLOP for Adyn x and 196 instead of 276 for AmBKt p.  `achieved` keeps SURVEY 8d's ALGORITHMIC
    # (dense) FLOP count; `executed_*` is what the FMA pipe really performs.
    exec_flop = FLOP_PER_ITER["q"] - (9 * (276 - 28) + 9 * (276 - 196) if stats.get("pattern") == 1 else 0)
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_fp32_max, "unit": "TFLOP/s",
                "frac": achieved / peak_fp32_max, "traffic": traffic,
                "structure_pattern": stats.get("pattern"), "executed_flop_per_iteration": exec_flop,
                "schedule": ("longest-expected-first: per-instance key kernel + CUB radix sort ahead of the solver kernel, on the same "
                             "stream, inside the timed region and inside kernel_ms_per_launch") if stats.get("scheduled") else "index order",
                "executed_tflops": stats["iterations"] * exec_flop / (kernel_ms * 1e-3) / 1e12,
                "executed_frac": stats["iterations"] * exec_flop / (kernel_ms * 1e-3) / 1e12 / peak_fp32_max,
                "peak_source": "SMs x 128 FMA lanes x 2 x clocks.max.sm (%d SMs, %d MHz); at the median clock under load "
                               "(%s MHz) the peak is %.1f TFLOP/s -> frac %.3f" % (prop.multi_processor_count,
                                                                                 clocks["sm_max_mhz"] or 1965, sm_mhz,
                                                                                 peak_fp32, achieved / peak_fp32),
                "kernel": "tmpc::admm_kernel_f32<12,4,10,256,%s,cold,TMEM,%s>" % (args.policy.upper(), "PatQuadrotor" if stats.get("pattern") == 1 else "PatDense"), "kernel_ms_per_launch": kernel_ms,
                "algorithmic_flop_per_iteration": FLOP_PER_ITER["q"], "iterations_per_launch": stats["iterations"],
                "fp32_instr_slot_util": (stats["iterations"] * exec_flop) / (kernel_ms * 1e-3) /
                                        (prop.multi_processor_count * 128 * sm_mhz * 1e6) if args.policy == "parity" else None,
                "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": hbm_peak, "frac": hbm_ach / hbm_peak,
                        "algorithmic_bytes_per_solve": BYTES_PER_SOLVE["q"],
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}}

    # ---- end to end through the C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        hx0 = torch.from_numpy(x0_np).pin_memory()
        hxr = torch.from_numpy(xref_np).pin_memory()
        hx = torch.empty((B, prob.N, prob.nx), dtype=torch.float32).pin_memory()
        hu = torch.empty((B, prob.N - 1, prob.nu), dtype=torch.float32).pin_memory()
        hit = torch.empty(B, dtype=torch.int32).pin_memory()
        hst = torch.empty(B, dtype=torch.int32).pin_memory()
        hrs = torch.empty((B, 4), dtype=torch.float32).pin_memory()

        def estep():
            solver.solve_raw(B, hx0, hxr, True, capi.TMPC_MEM_HOST, hx, hu, hit, hst, hrs)

        for _ in range(2):
            estep()
        barrier()
        ta = time.perf_counter()
        esteps = max(2, min(args.steps, 5))
        for _ in range(esteps):
            estep()                  # synchronous: returns when the outputs are in host memory
        barrier()
        tb = time.perf_counter()
        et = torch.tensor([tb - ta], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        assert int(hit.sum()) == int(it.sum().item()), "host-path results differ from device-path results"
        # the PCIe floor of this step: the same D2H bytes as one plain pinned-memory copy (explains e2e vs value)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(3):
            hx.copy_(x, non_blocking=True); hu.copy_(u, non_blocking=True); hit.copy_(it, non_blocking=True)
            hst.copy_(st, non_blocking=True); hrs.copy_(rs, non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize()
        d2h_ms = c0.elapsed_time(c1) / 3
        d2h_bytes = int(B * (480 + 144 + 4 + 4 + 16))
        # the same call when the caller only needs the controls (x = NULL: u, iter, status, resid come back: 168 B/solve)
        def cstep():
            solver.solve_raw(B, hx0, hxr, True, capi.TMPC_MEM_HOST, None, hu, hit, hst, hrs)
        cstep()
        barrier()
        tc0 = time.perf_counter()
        for _ in range(esteps):
            cstep()
        barrier()
        ct = torch.tensor([time.perf_counter() - tc0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        e2e = {"value": total_inst * esteps / float(et[0]), "unit": "solves/s",
               "controls_only": {"value": total_inst * esteps / float(ct[0]), "d2h_bytes_per_step": int(B * (144 + 4 + 4 + 16)),
                                 "what": "same call with x = NULL (u, iter, status, resid returned)"},
               "h2d_bytes_per_step": int(B * 48 + 480), "d2h_bytes_per_step": d2h_bytes,
               "ms_per_step": 1e3 * float(et[0]) / esteps,
               "pcie_floor": {"d2h_ms_plain_copy": d2h_ms, "d2h_gbs": d2h_bytes / (d2h_ms * 1e-3) / 1e9,
                              "solves_per_s_if_only_d2h": B / (d2h_ms * 1e-3) * world},
               "steps": esteps, "how": "tmpc_solve(TMPC_MEM_HOST) on pinned host buffers: H2D, one persistent-kernel launch, D2H of 65,536-instance chunks gated by in-kernel completion counters (cuStreamWaitValue32)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_reference(pkg, args, prob)

    if rank == 0:
        out = {"metric": "batched MPC solves/sec (quadrotor nx12 nu4 N10)", "value": value, "unit": "solves/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": config, "clocks": clocks, "gpu_launches": launches,
               "iters_per_s": total_iters * args.steps / (ms * 1e-3), "mean_iters_per_solve": total_iters / total_inst,
               "solved_frac": total_solved / total_inst,
               "iter_hist_le20_le40_le60_le80_lt100_eq100": [float(hist[1:21].sum()), float(hist[21:41].sum()),
                                                              float(hist[41:61].sum()), float(hist[61:81].sum()),
                                                              float(hist[81:100].sum()), float(hist[100:].sum())],
               "lane_trips_per_iteration": stats["trips"] / max(stats["iterations"], 1),
               "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu}
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
