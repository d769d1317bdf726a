#!/usr/bin/env python3
"""Benchmark of the batched TinyMPC ADMM hot path (BASELINE.json metric: batched MPC solves/sec, quadrotor
nx12 nu4 N10; % of the FP32 roofline).

  python bench.py --gpus N --steps K --warmup W            our arm  (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            the reference's own CPU tiny_solve on the host cores
  python bench.py --impl multi ...                         (hand-run) ONE process, every visible GPU through tmpc_multi

A "step" = one tmpc_solve over one batch: configs[1] of BASELINE.json -- 1,048,576 quadrotor_hovering instances
with seeded random initial states (mult 0.25), one shared cache and Xref, cold start, per GPU (weak
scaling: rank r owns instance indices [r*B, (r+1)*B)).  `value` is whole-job solves/s with x0 resident in HBM
and x/u/iter/status/resid written to HBM; `e2e` is the same through the C ABI with pinned HOST buffers
(H2D of x0 and D2H of every output inside the timed region); `e2e.u0_only` is that call with the controls-only
output mask (u(:,0) + iter + status: what an MPC loop consumes, quadrotor_hovering.cpp:110).
The same JSON line carries `configs`: BASELINE configs 3, 4, 5, the cartpole closed loop and the dense (non-specialised)
instance of the headline kernel, each device-timed with its own roofline and an oracle check of a 10,240-instance prefix
(the oracle is the CHECKER there, outside every timed region).  PyTorch is used for device buffers, events,
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

FLOP_PER_ITER = {"q": 11918, "c": 1771, "l": 344058}     # 2*MAC + ELT
BYTES_PER_SOLVE = {"q": 680, "q_track": 1160, "c": 220, "l_warm": 128 + 6400 + 1568 + 8 + 2 * 17504}
REF_SAMPLE = 131072        # instances per step of the reference arm (a bounded sample of the 1,048,576-instance step)
CHECK_PREFIX = 10240       # instances of every config re-solved by the oracle and compared bit for bit


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "multi"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="instances per GPU per step")
    ap.add_argument("--mult", type=float, default=0.25, help="initial-state spread (SURVEY 8d)")
    ap.add_argument("--policy", default="parity", choices=["parity", "fast"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` object (configs 3/4/5, closed loop, dense instance)")
    ap.add_argument("--config-scale", type=int, default=1, help="divide the batch of every entry of `configs` (smoke runs)")
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


def reference_lib():
    """The reference's own tiny_solve (oracle/_ref, compiled from /root/reference) -- or, if absent, the plain-C port."""
    from oracle.pyoracle import OracleLib, RefLib
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    for cfg, need, desc in (("q_f32_v3", ("avx2", "fma"), "-O3 -mavx2 -mfma"), ("q_f32", (), "-O3 (SSE2)")):
        if RefLib.available(cfg) and all(n in flags for n in need):
            return "reference", RefLib(cfg), "oracle/_ref/libref_%s.so: reference admm.cpp + vendored Eigen, g++ %s" % (cfg, desc)
    return "port", OracleLib(), "oracle/libtinympc_oracle.so (plain-C restatement, gcc -O2)"


def cpu_run(kind, lib, prob, x0, xref, cores):
    t = time.perf_counter()
    r = lib.solve_batch(prob, x0, xref, nthreads=cores) if kind == "reference" else lib.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=cores)
    return time.perf_counter() - t, r


def cpu_reference(pkg, args, prob):
    """cpu_baseline of our arm: ONE bounded sample (about --cpu-seconds of CPU work) of the same workload on all host cores."""
    cores = os.cpu_count() or 1
    kind, lib, build = reference_lib()
    W = pkg.workloads
    n = 4000 * cores
    dt, r = cpu_run(kind, lib, prob, *W.quadrotor_hover_batch(0, n, mult=args.mult), cores)     # calibration, warms the threads
    n = int(max(n, min(n * args.cpu_seconds / max(dt, 1e-3), 4_000_000)))
    dt, r = cpu_run(kind, lib, prob, *W.quadrotor_hover_batch(0, n, mult=args.mult), cores)
    return {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": kind,
            "sample": "first %d instances of the same seeded workload (mult %.2f), %.1f s, %s, float" % (n, args.mult, dt, build),
            "iters_per_s": float(r.iter.sum()) / dt, "mean_iters": float(r.iter.mean()), "seconds": dt}


def reference_arm(pkg, args, prob, config):
    """--impl reference: W warm-up + EXACTLY K timed steps; a step = the reference's tiny_solve looping over a bounded sample
    (REF_SAMPLE consecutive instances of the 1,048,576-instance workload step, a different slice every step) on all host
    cores.  ms_per_step is the measured time of such a step, so steps x ms_per_step is the timed region."""
    cores = os.cpu_count() or 1
    kind, lib, build = reference_lib()
    W = pkg.workloads
    n = REF_SAMPLE
    steps, warm = max(1, args.steps), max(1, args.warmup)
    slices = [W.quadrotor_hover_batch(k * n, (k + 1) * n, mult=args.mult) for k in range(min(8, steps + warm))]
    for k in range(warm):
        cpu_run(kind, lib, prob, *slices[k % len(slices)], cores)
    iters = 0
    t0 = time.perf_counter()
    for k in range(steps):
        _, r = cpu_run(kind, lib, prob, *slices[(warm + k) % len(slices)], cores)
        iters += int(r.iter.sum())
    dt = time.perf_counter() - t0
    v = steps * n / dt
    sample = ("each step = tiny_solve over %d consecutive instances of the seeded 1,048,576-instance workload step (mult %.2f), "
              "%d steps in %.1f s, %s, float, %d threads" % (n, args.mult, steps, dt, build, cores))
    cfg = dict(config)
    cfg["reference_step"] = "bounded sample: %d of the %d instances of one workload step" % (n, args.batch)
    return {"impl": "reference", "metric": "batched MPC solves/sec (quadrotor nx12 nu4 N10)", "value": v,
            "unit": "solves/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "iters_per_s": iters / dt,
            "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything any library prints to fd 1 during the run
    (torch.distributed's NCCL banner, for one) has been routed to stderr by main()."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------------------
# device-resident measurement of one workload (used by the headline and by every entry of `configs`)
# ------------------------------------------------------------------------------------------------------------------
class DeviceRun:
    def __init__(self, torch, pkg, prob, dev, local, policy, x0_np, xref_np, warm=False, dtype=np.float32):
        self.torch, self.capi, self.prob, self.dev = torch, pkg.capi, prob, dev
        self.B = x0_np.shape[0]
        self.shared = xref_np.ndim == 2
        self.solver = pkg.capi.Solver(prob, dtype=dtype, policy=policy, device=local)
        B = self.B
        self.x0 = torch.from_numpy(x0_np.astype(dtype)).to(dev)
        self.xref = torch.from_numpy(xref_np.astype(dtype)).to(dev)
        f32, i32 = (torch.float32 if np.dtype(dtype) == np.float32 else torch.float64), torch.int32
        self.x = torch.empty((B, prob.N, prob.nx), dtype=f32, device=dev)
        self.u = torch.empty((B, prob.N - 1, prob.nu), dtype=f32, device=dev)
        self.it = torch.empty(B, dtype=i32, device=dev)
        self.st = torch.empty(B, dtype=i32, device=dev)
        self.rs = torch.empty((B, 4), dtype=f32, device=dev)
        self.warm = None
        if warm:
            zu = lambda: torch.zeros((B, prob.N - 1, prob.nu), dtype=f32, device=dev)
            zx = lambda: torch.zeros((B, prob.N, prob.nx), dtype=f32, device=dev)
            self.warm = {"d": zu(), "y": zu(), "z": zu(), "g": zx(), "v": zx()}
        self.stream = torch.cuda.Stream(device=dev)

    def step(self):
        self.solver.solve_raw(self.B, self.x0, self.xref, self.shared, self.capi.TMPC_MEM_DEVICE, self.x, self.u, self.it, self.st,
                              self.rs, warm=self.warm, stream=self.stream.cuda_stream)

    def timed(self, steps, warmup, barrier, before_each=None):
        """W untimed + K timed steps bracketed by barrier + synchronize; returns (ms for the K steps, stats of the last)."""
        torch = self.torch
        for _ in range(warmup):
            if before_each:
                before_each()
            self.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if before_each is None:
            e0.record(self.stream)
            for _ in range(steps):
                self.step()
            e1.record(self.stream)
            barrier()
            ms = e0.elapsed_time(e1)
        else:   # state restored between steps (warm re-solve): time each step on its own, restoration outside
            ms = 0.0
            for _ in range(steps):
                before_each()
                torch.cuda.synchronize()
                e0.record(self.stream)
                self.step()
                e1.record(self.stream)
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            barrier()
        return ms, self.solver.stats()

    def close(self):
        self.solver.close()


def oracle_check(prob, x0, xref, out, n, warm_in=None, what=("iter", "status", "x", "u"), dtype=np.float32):
    """Bit-for-bit comparison of the first n instances with the CPU oracle (checker only; never inside a timed region)."""
    from oracle.pyoracle import OracleLib
    n = min(n, x0.shape[0])
    xr = xref if xref.ndim == 2 else xref[:n]
    w = None if warm_in is None else {k: v[:n] for k, v in warm_in.items()}
    t = time.perf_counter()
    ref = OracleLib().solve_batch(prob, x0[:n], xr, dtype=dtype, warm=w, nthreads=os.cpu_count() or 1)
    bad = {k: int((np.asarray(out[k][:n]) != getattr(ref, k)).sum()) for k in what}
    return {"instances": n, "bit_exact": all(v == 0 for v in bad.values()), "mismatching_elements": bad, "compared": list(what),
            "oracle_seconds": time.perf_counter() - t}


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
        emit(json.dumps(reference_arm(pkg, args, prob, config)))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if args.impl == "multi":
        return multi_arm(torch, pkg, args, prob, config)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi = pkg.capi
    B = args.batch
    b0, b1 = pkg.sharding.shard_range(rank, world, per_rank=B)
    x0_np, xref_np = pkg.workloads.quadrotor_hover_batch(b0, b1, mult=args.mult)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    # ---- headline: device-resident, W warm-up + K timed steps
    run = DeviceRun(torch, pkg, prob, dev, local, args.policy, x0_np, xref_np)
    torch.cuda.set_stream(run.stream)
    solver, x, u, it, st, rs, stream = run.solver, run.x, run.u, run.it, run.st, run.rs, run.stream
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        run.step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    t0 = time.time()
    ms, stats = run.timed(args.steps, 0, barrier)
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    kernel_ms = stats["kernel_ms"]       # CUDA-event time of the LAST step's launch (recorded on the launching stream)
    launches = stats["launches"] * args.steps

    # whole-job numbers: max time over ranks, totals over ranks (NCCL only for this statistics gather)
    vec = pkg.sharding.local_stats(it.cpu().numpy(), st.cpu().numpy(), prob.max_iter)
    assert vec[0] == stats["iterations"] and vec[1] == stats["solved"], "kernel statistics disagree with the outputs"
    vec, tmax = pkg.sharding.gather_stats(vec, [ms, kernel_ms], dist if world > 1 else None, dev)
    ms, kernel_ms = float(tmax[0]), float(tmax[1])
    total_iters, total_solved, total_inst = float(vec[0]), float(vec[1]), float(vec[2])
    hist = vec[3:]
    value = total_inst * args.steps / (ms * 1e-3)

    # ---- cross-rank determinism on hardware: every rank solves the SAME probe slice; checksums must agree (MIN == MAX)
    probe = None
    if world > 1:
        pn = 16384
        px0, pxr = pkg.workloads.quadrotor_hover_batch(0, pn, mult=args.mult)
        prun = DeviceRun(torch, pkg, prob, dev, local, args.policy, px0, pxr)
        prun.step()
        torch.cuda.synchronize()
        sums = torch.stack([prun.it.to(torch.int64).sum(), prun.st.to(torch.int64).sum(),
                            prun.u.view(torch.int32).to(torch.int64).sum(), prun.x.view(torch.int32).to(torch.int64).sum()])
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        probe = {"instances": pn, "identical_on_every_rank": bool((lo == hi).all().item()),
                 "what": "sum(iter), sum(status), integer sums of the bit patterns of u and x of a common 16,384-instance slice"}
        prun.close()

    # ---- roofline of the one kernel (per GPU, per launch; kernel time by CUDA events inside the library,
    #      recorded on the launching stream around the launch)
    prop = torch.cuda.get_device_properties(local)
    sm_mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965
    sm_max = clocks["sm_max_mhz"] or 1965
    peak_fp32 = prop.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12          # TFLOP/s at the clock seen under load
    peak_fp32_max = prop.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12
    flops = stats["iterations"] * FLOP_PER_ITER["q"]
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"
    hbm_ach = stats["instances"] * BYTES_PER_SOLVE["q"] / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:   # DRAM bytes (read + write) of this kernel at this workload: STATIC, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        if tj.get("batch") == args.batch and tj.get("policy") == args.policy and abs(tj.get("mult", -1) - args.mult) < 1e-9:
            traffic = tj["dram_bytes_per_launch"]
            traffic_src = "static: %s (not measured in this run)" % tj.get("source", "profiles/kernel_traffic.json")
    except Exception:
        pass
    # With the quadrotor structure specialisation (stats.pattern == 1) the kernel drops the terms whose coefficient is an
    # exact zero (118 of Adyn's 144, 40 of AmBKt's 144) and the multiplies by exact ones (Adyn's diagonal): per stage
    # 28 instead of 276 FLOP for Adyn x and 196 instead of 276 for AmBKt p.  `achieved` keeps SURVEY 8d's ALGORITHMIC
    # (dense) FLOP count; `executed_*` is what the FMA pipe really performs.
    def exec_flop_of(pattern):
        return FLOP_PER_ITER["q"] - (9 * (276 - 28) + 9 * (276 - 196) if pattern == 1 else 0)
    exec_flop = exec_flop_of(stats.get("pattern"))
    sched = {0: "index order", 1: ("longest-expected-first: per-instance key kernel + CUB radix sort ahead of the solver kernel, on the same "
                                   "stream, inside the timed region and inside kernel_ms_per_launch"), 2: "tail-sorted (host pipeline)"}
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_fp32_max, "unit": "TFLOP/s",
                "frac": achieved / peak_fp32_max, "traffic": traffic, "traffic_source": traffic_src,
                "structure_pattern": stats.get("pattern"), "executed_flop_per_iteration": exec_flop,
                "schedule": sched.get(stats.get("scheduled"), "?"),
                "executed_tflops": stats["iterations"] * exec_flop / (kernel_ms * 1e-3) / 1e12,
                "executed_frac": stats["iterations"] * exec_flop / (kernel_ms * 1e-3) / 1e12 / peak_fp32_max,
                "peak_source": "SMs x 128 FMA lanes x 2 x clocks.max.sm (%d SMs, %d MHz; MEASURED_PEAKS.json has no FP32 entry); at the "
                               "median clock under load (%s MHz) the peak is %.1f TFLOP/s -> frac %.3f"
                               % (prop.multi_processor_count, sm_max, sm_mhz, peak_fp32, achieved / peak_fp32),
                "kernel": "tmpc::admm_kernel_f32<12,4,10,256,%s,cold,TMEM,%s>" % (args.policy.upper(), "PatQuadrotor" if stats.get("pattern") == 1 else "PatDense"),
                "kernel_ms_per_launch": kernel_ms,
                "algorithmic_flop_per_iteration": FLOP_PER_ITER["q"], "iterations_per_launch": stats["iterations"],
                "fp32_instr_slot_util": (stats["iterations"] * exec_flop) / (kernel_ms * 1e-3) /
                                        (prop.multi_processor_count * 128 * sm_mhz * 1e6) if args.policy == "parity" else None,
                "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": hbm_peak, "frac": hbm_ach / hbm_peak,
                        "algorithmic_bytes_per_solve": BYTES_PER_SOLVE["q"], "peak_source": hbm_src}}

    # ---- end to end through the C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        hx0 = torch.from_numpy(x0_np).pin_memory()
        hxr = torch.from_numpy(xref_np).pin_memory()
        hx = torch.empty((B, prob.N, prob.nx), dtype=torch.float32).pin_memory()
        hu = torch.empty((B, prob.N - 1, prob.nu), dtype=torch.float32).pin_memory()
        hu0 = torch.empty((B, prob.nu), dtype=torch.float32).pin_memory()
        hit = torch.empty(B, dtype=torch.int32).pin_memory()
        hst = torch.empty(B, dtype=torch.int32).pin_memory()
        hrs = torch.empty((B, 4), dtype=torch.float32).pin_memory()
        esteps = max(2, min(args.steps, 5))

        def e2e_timed(fn):
            for _ in range(2):
                fn()
            barrier()
            ta = time.perf_counter()
            for _ in range(esteps):
                fn()                  # synchronous: returns when the outputs are in host memory
            barrier()
            return allmax(time.perf_counter() - ta), solver.stats()

        full_s, full_stats = e2e_timed(lambda: solver.solve_raw(B, hx0, hxr, True, capi.TMPC_MEM_HOST, hx, hu, hit, hst, hrs))
        assert int(hit.sum()) == int(it.sum().item()), "host-path results differ from device-path results"
        assert torch.equal(hu, u.cpu()), "host-path u differs from device-path u"
        u0_s, u0_stats = e2e_timed(lambda: solver.solve_raw(B, hx0, hxr, True, capi.TMPC_MEM_HOST, None, None, hit, hst, None, u0=hu0))
        assert int(hit.sum()) == int(it.sum().item()) and torch.equal(hu0, u[:, 0, :].cpu()), "controls-only results differ"
        # the PCIe floor of the full-output step: the same D2H bytes as one plain pinned-memory copy (explains e2e vs value)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        c0.record(stream)
        for _ in range(3):
            hx.copy_(x, non_blocking=True); hu.copy_(u, non_blocking=True); hit.copy_(it, non_blocking=True)
            hst.copy_(st, non_blocking=True); hrs.copy_(rs, non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize()
        d2h_ms = allmax(c0.elapsed_time(c1) / 3)
        d2h_bytes = int(B * (480 + 144 + 4 + 4 + 16))
        e2e = {"value": total_inst * esteps / full_s, "unit": "solves/s",
               "h2d_bytes_per_step": int(B * 48 + 480), "d2h_bytes_per_step": d2h_bytes,
               "ms_per_step": 1e3 * full_s / esteps, "kernel_ms_last_step": full_stats["kernel_ms"],
               "u0_only": {"value": total_inst * esteps / u0_s, "unit": "solves/s", "ms_per_step": 1e3 * u0_s / esteps,
                           "h2d_bytes_per_step": int(B * 48 + 480), "d2h_bytes_per_step": int(B * (16 + 4 + 4)),
                           "frac_of_device_timed_value": (total_inst * esteps / u0_s) / value,
                           "kernel_ms_last_step": u0_stats["kernel_ms"],
                           "lane_trips_per_iteration": u0_stats["trips"] / max(u0_stats["iterations"], 1),
                           "what": "same call with the controls-only output mask: x = u = NULL, u0 = u(:,0) + iter + status returned "
                                   "(what an MPC loop applies, quadrotor_hovering.cpp:110); the kernel skips the emission pass"},
               "pcie_floor": {"d2h_ms_plain_copy": d2h_ms, "d2h_gbs_per_rank": d2h_bytes / (d2h_ms * 1e-3) / 1e9,
                              "solves_per_s_if_only_d2h": B / (d2h_ms * 1e-3) * world,
                              "what": "the full-output step's D2H bytes as plain pinned copies on every rank at once (max over ranks)"},
               "schedule": sched.get(full_stats.get("scheduled"), "?"),
               "steps": esteps, "how": "tmpc_solve(TMPC_MEM_HOST) on pinned host buffers: H2D in chunks behind an in-kernel arrival gate, one "
                                       "persistent-kernel launch, D2H of 65,536-instance chunks gated by in-kernel completion counters "
                                       "(cuStreamWriteValue32 / cuStreamWaitValue32); leading quarter of the batch ranked and claimed last"}
        del hx, hu, hrs
    run.close()
    del run, x, u, rs
    torch.cuda.empty_cache()

    # ---- the same end-to-end step from ONE process over all N GPUs (tmpc_multi: what a C++ caller of tiny_solve_batch gets):
    #      rank 0 drives every device, the other ranks wait on the rendezvous store (a CPU wait: no NCCL kernel on their GPUs)
    if e2e is not None and world > 1:
        store = dist.distributed_c10d._get_default_store()
        barrier()
        if rank == 0:
            try:
                e2e["one_process"] = multi_measure(torch, pkg, args, prob, world)
            except Exception as ex:   # additional evidence only
                e2e["one_process"] = {"error": repr(ex)}
            store.set("tmpc_one_process_done", "1")
        else:
            store.wait(["tmpc_one_process_done"])
        barrier()

    # ---- the other BASELINE configs + closed loop + dense instance, device-timed, same one JSON line
    configs = None
    if not args.no_configs:
        configs = run_configs(torch, dist, pkg, args, dev, local, rank, world, barrier, allmax, allsum, peak_fp32_max, hbm_peak, hbm_src,
                              exec_flop_of)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(pkg, args, prob)

    if rank == 0:
        out = {"metric": "batched MPC solves/sec (quadrotor nx12 nu4 N10)", "value": value, "unit": "solves/s",
               "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": config, "clocks": clocks, "gpu_launches": launches,
               "iters_per_s": total_iters * args.steps / (ms * 1e-3), "mean_iters_per_solve": total_iters / total_inst,
               "solved_frac": total_solved / total_inst,
               "iter_hist_le20_le40_le60_le80_lt100_eq100": [float(hist[1:21].sum()), float(hist[21:41].sum()),
                                                              float(hist[41:61].sum()), float(hist[61:81].sum()),
                                                              float(hist[81:100].sum()), float(hist[100:].sum())],
               "lane_trips_per_iteration": stats["trips"] / max(stats["iterations"], 1),
               "roofline": roofline, "e2e": e2e, "cross_rank_probe": probe, "configs": configs, "cpu_baseline": cpu}
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_configs(torch, dist, pkg, args, dev, local, rank, world, barrier, allmax, allsum, peak_tf, hbm_peak, hbm_src, exec_flop_of):
    """BASELINE configs 3, 4, 5 at their named TOTAL sizes, strong-scaled over the ranks (rank r owns [T r / G, T (r+1) / G)),
    plus the cartpole closed loop (the regime BASELINE calls latency-bound) and the dense instance of the headline kernel.
    Device-resident, CUDA events on the launching stream, max over ranks; rank 0 re-solves a prefix of its shard with the
    oracle and compares bit for bit (outside the timed regions)."""
    W, P = pkg.workloads, pkg.problems
    K, WU = 5, 2
    sc = max(1, args.config_scale)
    out = {"steps": K, "warmup": WU, "scaling": "strong: every entry is the BASELINE total split over the ranks" if world > 1 else "single GPU",
           "timing": "CUDA events on the launching stream around K back-to-back steps after W warm-up steps, max over ranks; "
                     "kernel_ms_per_launch = the library's own events around the last launch",
           "l2": "every step reads or writes more than the 126 MB L2 per GPU (sizes in each entry)"}

    def entry(name, workload, shape, run, total, ms, st, bytes_per_solve, check, extra=None):
        iters = allsum(st["iterations"])
        kms = allmax(st["kernel_ms"])
        tf = st["iterations"] * FLOP_PER_ITER[shape] / (st["kernel_ms"] * 1e-3) / 1e12      # this rank's kernel
        gbs = st["instances"] * bytes_per_solve / (st["kernel_ms"] * 1e-3) / 1e9
        bound = "hbm" if gbs / hbm_peak > tf / peak_tf else "fp32"
        e = {"workload": workload, "instances_total": total, "instances_per_gpu": run.B, "value": total * K / (ms * 1e-3), "unit": "solves/s",
             "ms_per_step": ms / K, "kernel_ms_per_launch": kms, "iters_per_s": iters * K / (ms * 1e-3), "mean_iters_per_solve": iters / total,
             "solved_frac": allsum(st["solved"]) / total, "scheduled": st.get("scheduled"), "pattern": st.get("pattern"),
             "lane_trips_per_iteration": st["trips"] / max(st["iterations"], 1),
             "roofline": {"bound": bound, "achieved": gbs if bound == "hbm" else tf, "peak": hbm_peak if bound == "hbm" else peak_tf,
                          "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": gbs / hbm_peak if bound == "hbm" else tf / peak_tf,
                          "fp32": {"achieved_tflops": tf, "peak": peak_tf, "frac": tf / peak_tf, "algorithmic_flop_per_iteration": FLOP_PER_ITER[shape]},
                          "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak, "algorithmic_bytes_per_solve": bytes_per_solve,
                                  "peak_source": hbm_src},
                          "traffic": None},
             "oracle_check": check}
        if extra:
            e.update(extra)
        out[name] = e
        log("configs.%s: %.3e solves/s, %.2f ms/step, frac %.3f (%s)" % (name, e["value"], e["ms_per_step"], e["roofline"]["frac"], bound))

    def outputs_np(run):
        return {"iter": run.it.cpu().numpy(), "status": run.st.cpu().numpy(), "x": run.x[:CHECK_PREFIX].cpu().numpy(),
                "u": run.u[:CHECK_PREFIX].cpu().numpy()}

    quad = P.quadrotor(20)
    # ---- config 3: quadrotor_tracking, per-instance reference windows, 4,194,304 instances in total
    T = (1 << 22) // sc
    b0, b1 = pkg.sharding.shard_range(rank, world, total=T)
    x0, xref = W.quadrotor_tracking_batch(b0, b1)
    run = DeviceRun(torch, pkg, quad, dev, local, args.policy, x0, xref)
    ms, st = run.timed(K, WU, barrier)
    ms = allmax(ms)
    chk = oracle_check(quad, x0, xref, outputs_np(run), CHECK_PREFIX) if rank == 0 else None
    entry("config3_tracking", "quadrotor_tracking (BASELINE configs[2]): per-instance Xref windows k_b = b mod 290 of the y-axis-line table, "
          "x0 = window start + 0.1 scale noise, cold start; %d B of per-instance inputs per solve" % (48 + 480), "q", run, T, ms, st,
          BYTES_PER_SOLVE["q_track"], chk)
    run.close(); del run; torch.cuda.empty_cache()

    # ---- dense instance of the headline workload (what any model that is not the shipped quadrotor gets)
    Bd = args.batch // sc
    b0, b1 = pkg.sharding.shard_range(rank, world, per_rank=Bd)
    x0, xref = W.quadrotor_hover_batch(b0, b1, mult=args.mult)
    os.environ["TMPC_DENSE"] = "1"
    try:
        run = DeviceRun(torch, pkg, quad, dev, local, args.policy, x0, xref)
    finally:
        del os.environ["TMPC_DENSE"]
    ms, st = run.timed(K, WU, barrier)
    ms = allmax(ms)
    chk = oracle_check(quad, x0, xref, outputs_np(run), CHECK_PREFIX) if rank == 0 else None
    entry("headline_dense_instance", "the headline workload (configs[1], %d instances per GPU, weak) on the DENSE kernel instance: no "
          "model-structure specialisation (TMPC_DENSE=1), every one of the 11,918 FLOP per iteration executed" % Bd, "q", run, Bd * world, ms, st,
          BYTES_PER_SOLVE["q"], chk, {"executed_flop_per_iteration": exec_flop_of(st.get("pattern"))})
    run.close(); del run; torch.cuda.empty_cache()

    # ---- the examples' closed loop (quadrotor_hovering.cpp:90-114) on the device: 10 MPC steps per instance from the cold start
    #      (reset duals, warm d / v / z, plant step x <- Adyn x + Bdyn u0); fp32 12/4/10 runs it as ONE persistent launch
    try:
        Bh = args.batch // sc
        b0, b1 = pkg.sharding.shard_range(rank, world, per_rank=Bh)
        x0, xref = W.quadrotor_hover_batch(b0, b1, mult=args.mult)
        s = pkg.capi.Solver(quad, dtype=np.float32, policy=args.policy, device=local)
        bt = pkg.capi.Batch(s, Bh)
        steps_h = 10
        ith = torch.empty((steps_h, Bh), dtype=torch.int32, device=dev)
        sth = torch.empty((steps_h, Bh), dtype=torch.int32, device=dev)
        u0h = torch.empty((steps_h, Bh, 4), dtype=torch.float32, device=dev)
        x0h = torch.empty((steps_h + 1, Bh, 12), dtype=torch.float32, device=dev)
        reps_ms = []
        for rep in range(3):     # (each repetition restarts the loop from the cold workspace; the first one also allocates)
            bt.reset(); bt.set_x0(x0); bt.set_xref(xref)
            barrier()
            s._check(s.lib.tmpc_batch_rollout(bt._b, steps_h, 1, x0h.data_ptr(), u0h.data_ptr(), ith.data_ptr(), sth.data_ptr(),
                                              pkg.capi.TMPC_MEM_DEVICE), "rollout")
            reps_ms.append(allmax(bt.last_rollout_ms()))
        ms_h = min(reps_ms[1:])
        launches_h = s.stats()["launches"]
        iters_h = allsum(float(ith.sum().item()))
        chk = None
        if rank == 0:
            from oracle.pyoracle import OracleLib
            orc = OracleLib()
            t0c = time.perf_counter()
            n = min(CHECK_PREFIX // 4, Bh)
            xc, warm, bad = x0[:n].copy(), None, {"iter": 0, "status": 0, "u0": 0, "plant_state": 0}
            for k in range(steps_h):
                r = orc.solve_batch(quad, xc, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=os.cpu_count() or 1)
                bad["iter"] += int((ith[k, :n].cpu().numpy() != r.iter).sum())
                bad["status"] += int((sth[k, :n].cpu().numpy() != r.status).sum())
                bad["u0"] += int((u0h[k, :n].cpu().numpy() != r.u[:, 0, :]).sum())
                xc = orc.plant_step(quad, xc, r.u[:, 0, :], dtype=np.float32)
                bad["plant_state"] += int((x0h[k + 1, :n].cpu().numpy() != xc).sum())
                warm = {q: r.state[q].copy() for q in ("d", "y", "g", "v", "z")}
                warm["y"][:] = 0
                warm["g"][:] = 0
            chk = {"instances": n, "mpc_steps": steps_h, "bit_exact": all(v == 0 for v in bad.values()), "mismatching_elements": bad,
                   "compared": list(bad), "oracle_seconds": time.perf_counter() - t0c,
                   "how": "the oracle's closed loop (solve, its own plant step, duals reset) for the first %d instances, every step" % n}
        tf = iters_h / world * FLOP_PER_ITER["q"] / (ms_h * 1e-3) / 1e12
        out["closed_loop_rollout_hover"] = {
            "workload": "the hovering example's closed loop (quadrotor_hovering.cpp:90-114) for every instance of the headline batch, on the device: "
                        "%d MPC steps from the cold start, duals reset every step, warm d / v / z, plant step; histories of x, u0, iter, status "
                        "written to HBM; tmpc_batch_rollout" % steps_h,
            "instances_total": Bh * world, "instances_per_gpu": Bh, "mpc_steps": steps_h, "value": Bh * world * steps_h / (ms_h * 1e-3),
            "unit": "MPC steps/s (= solves/s)", "ms_per_rollout": ms_h, "ms_per_rollout_all_repetitions": reps_ms, "ms_per_mpc_step": ms_h / steps_h, "launches_per_rollout": launches_h,
            "fused": os.environ.get("TMPC_ROLL", "1") != "0",
            "iters_per_s": iters_h / (ms_h * 1e-3), "mean_iters_per_solve": iters_h / (Bh * world * steps_h),
            "roofline": {"bound": "fp32", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "algorithmic_flop_per_iteration": FLOP_PER_ITER["q"], "traffic": None},
            "oracle_check": chk}
        log("configs.closed_loop_rollout_hover: %.3e MPC steps/s, %.2f ms per %d-step rollout, frac %.3f" %
            (out["closed_loop_rollout_hover"]["value"], ms_h, steps_h, tf / peak_tf))
        bt.close(); s.close(); del ith, sth, u0h, x0h
    except Exception as e:   # additional evidence: a failure here must not void the headline line
        out["closed_loop_rollout_hover"] = {"error": repr(e)}
    torch.cuda.empty_cache()

    # ---- the reference's SHIPPED scalar type (glob_opts.hpp:3 `typedef double tinytype`): the headline workload in fp64, a quarter
    #      of the batch per GPU (half the instances per SM of the fp32 kernel, a quarter of its arithmetic rate)
    Bf = max(args.batch // 4 // sc, 1024)
    b0, b1 = pkg.sharding.shard_range(rank, world, per_rank=Bf)
    x0, xref = W.quadrotor_hover_batch(b0, b1, mult=args.mult)
    run = DeviceRun(torch, pkg, quad, dev, local, args.policy, x0, xref, dtype=np.float64)
    ms, st = run.timed(K, WU, barrier)
    ms = allmax(ms)
    chk = oracle_check(quad, x0, xref, outputs_np(run), CHECK_PREFIX, dtype=np.float64) if rank == 0 else None
    prop = torch.cuda.get_device_properties(local)
    peak64 = prop.multi_processor_count * 64 * 2 * 1965e6 / 1e12
    tf64 = st["iterations"] * FLOP_PER_ITER["q"] / (st["kernel_ms"] * 1e-3) / 1e12
    out["headline_fp64"] = {
        "workload": "the headline workload in double precision (the reference's shipped tinytype), %d instances per GPU, PARITY; two lanes per instance, model in shared memory (tmpc_kernel_f64p.cuh)" % Bf,
        "instances_total": Bf * world, "instances_per_gpu": Bf, "value": Bf * world * K / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms / K,
        "kernel_ms_per_launch": allmax(st["kernel_ms"]), "iters_per_s": allsum(st["iterations"]) * K / (ms * 1e-3),
        "mean_iters_per_solve": allsum(st["iterations"]) / (Bf * world), "dtype": "f64",
        "roofline": {"bound": "fp64", "achieved": tf64, "peak": peak64, "unit": "TFLOP/s", "frac": tf64 / peak64,
                     "peak_source": "SMs x 64 FP64 lanes x 2 x 1965 MHz (nominal; MEASURED_PEAKS.json has no FP64 entry)",
                     "algorithmic_flop_per_iteration": FLOP_PER_ITER["q"], "traffic": None},
        "oracle_check": chk}
    log("configs.headline_fp64: %.3e solves/s, frac %.3f of FP64" % (out["headline_fp64"]["value"], tf64 / peak64))
    run.close(); del run; torch.cuda.empty_cache()

    # ---- per-instance SYSTEMS (every instance its own Adyn, Bdyn, rho and cache), 262,144 in total
    T = (1 << 18) // sc
    b0, b1 = pkg.sharding.shard_range(rank, world, total=T)
    rng = np.random.default_rng(1)
    off = ~np.eye(12, dtype=bool)
    sa, sb_, sr = rng.uniform(-1, 1, (T, 1)), rng.uniform(-1, 1, (T, 1, 1)), rng.uniform(-1, 1, T)
    A = np.repeat(quad.Adyn[None], b1 - b0, 0).copy()
    A[:, off] *= (1.0 + 0.2 * sa[b0:b1])
    Bm = quad.Bdyn[None] * (1.0 + 0.3 * sb_[b0:b1])
    Qs, Rs = np.repeat(quad.Q[None], b1 - b0, 0), np.repeat(quad.R[None], b1 - b0, 0)
    rhos = 5.0 * (1.0 + 0.4 * sr[b0:b1])
    x0, xref = W.quadrotor_hover_batch(b0, b1, mult=args.mult)
    run = DeviceRun(torch, pkg, quad, dev, local, args.policy, x0, xref)
    sysobj = pkg.capi.Systems(run.solver, A, Bm, Qs, Rs, rhos)
    run.step = lambda: sysobj.solve_raw(run.x0, run.xref, True, run.x, run.u, run.it, run.st, run.rs, stream=run.stream.cuda_stream)
    ms, st = run.timed(K, WU, barrier)
    ms = allmax(ms)
    chk = None
    if rank == 0:   # the oracle solves a sample of the systems one by one, each with the cache the device computed for it
        import copy
        from oracle.pyoracle import OracleLib
        orc = OracleLib()
        t0c = time.perf_counter()
        Kc, Pc, Qic, Mc = sysobj.get("Kinf"), sysobj.get("Pinf"), sysobj.get("Quu_inv"), sysobj.get("AmBKt")
        o = outputs_np(run)
        bad = {"iter": 0, "status": 0, "x": 0, "u": 0}
        sample = list(range(0, min(b1 - b0, CHECK_PREFIX), max(1, min(b1 - b0, CHECK_PREFIX) // 256)))
        for j in sample:
            pj = copy.deepcopy(quad)
            pj.Adyn, pj.Bdyn, pj.rho = A[j].astype(np.float32).astype(np.float64), Bm[j].astype(np.float32).astype(np.float64), float(np.float32(rhos[j]))
            pj.Kinf, pj.Pinf, pj.Quu_inv, pj.AmBKt = (Kc[j].astype(np.float64), Pc[j].astype(np.float64), Qic[j].astype(np.float64), Mc[j].astype(np.float64))
            ref = orc.solve_batch(pj, x0[j:j + 1], xref, dtype=np.float32, nthreads=1)
            for k in bad:
                bad[k] += int((np.asarray(o[k][j:j + 1]) != getattr(ref, k)).sum())
        chk = {"instances": len(sample), "bit_exact": all(v == 0 for v in bad.values()), "mismatching_elements": bad, "compared": list(bad),
               "oracle_seconds": time.perf_counter() - t0c, "how": "every %d-th system of the first %d, each solved by the oracle with its own model and the cache "
               "the batched device precompute produced for it" % (max(1, min(b1 - b0, CHECK_PREFIX) // 256), min(b1 - b0, CHECK_PREFIX))}
    entry("per_instance_systems", "SURVEY 8f row 1: %d perturbed quadrotor systems in total (each instance its own Adyn, Bdyn, rho, cache from the batched "
          "device precompute), hover batch, cold, tmpc_solve_systems; two lanes per instance, coefficients streamed from tensor memory" % T,
          "q", run, T, ms, st, BYTES_PER_SOLVE["q"] + 4 * 656, chk, {"kernel": os.environ.get("TMPC_KERNEL", "sys lane pairs (default)")})
    sysobj.close(); run.close(); del run, sysobj; torch.cuda.empty_cache()

    # ---- config 4: cartpole 4/1/10, 16,777,216 instances in total, cold start
    cart = P.cartpole()
    T = (1 << 24) // sc
    b0, b1 = pkg.sharding.shard_range(rank, world, total=T)
    x0, xref = W.cartpole_batch(b0, b1)
    run = DeviceRun(torch, pkg, cart, dev, local, args.policy, x0, xref)
    ms, st = run.timed(K, WU, barrier)
    ms = allmax(ms)
    chk = oracle_check(cart, x0, xref, outputs_np(run), CHECK_PREFIX) if rank == 0 else None
    entry("config4_cartpole_cold", "codegen_cartpole 4/1/10 (BASELINE configs[3]): cache by the codegen recursion, x0 = {0.5,0.5,0.2,0.5} U(-1,1), "
          "Xref = 0, cold start (SURVEY 8d recipe)", "c", run, T, ms, st, BYTES_PER_SOLVE["c"], chk)
    run.close(); del run; torch.cuda.empty_cache()
    # the closed loop the reference actually runs with this model (codegen_cartpole.cpp:75-122): measurement -> reset duals ->
    # warm-started tiny_solve (1-4 iterations) -> plant step, state resident in HBM between steps (tmpc_batch_rollout)
    try:
        s = pkg.capi.Solver(cart, dtype=np.float32, policy=args.policy, device=local)
        Bc = x0.shape[0]
        bt = pkg.capi.Batch(s, Bc)
        # 0.2 x the cold recipe's spread: perturbations of the size the reference's own loop starts from (x0 = (0, 0, 0.1, 0),
        # codegen_cartpole.cpp:87) -- every instance stays feasible and a step takes 1-4 iterations, like the reference's run
        # (SURVEY G5); at the cold recipe's spread 31-48 % of the instances sit at max_iter for the whole loop
        bt.set_x0((0.2 * x0).astype(np.float32)); bt.set_xref(xref)
        steps_cl = 8
        ith = torch.empty((steps_cl, Bc), dtype=torch.int32, device=dev)
        s._check(s.lib.tmpc_batch_rollout(bt._b, 4, 1, None, None, None, None, pkg.capi.TMPC_MEM_DEVICE), "rollout")     # leave the cold steps behind
        barrier()
        s._check(s.lib.tmpc_batch_rollout(bt._b, steps_cl, 1, None, None, ith.data_ptr(), None, pkg.capi.TMPC_MEM_DEVICE), "rollout")
        ms_cl = allmax(bt.last_rollout_ms())
        iters_cl = allsum(float(ith.sum().item()))
        hist = torch.bincount(ith.flatten().to(torch.int64), minlength=8)[:8].cpu().numpy().tolist()
        chk_cl = None
        if rank == 0:     # the oracle runs the same 4 + STEPS steps for a prefix: iteration counts of the timed steps and the final plant state
            from oracle.pyoracle import OracleLib
            orc = OracleLib()
            t0c = time.perf_counter()
            n = min(CHECK_PREFIX, Bc)
            xc, warm, bad = (0.2 * x0[:n]).astype(np.float32), None, {"iter": 0, "plant_state": 0}
            ith_h = ith[:, :n].cpu().numpy()
            for k in range(4 + steps_cl):
                r = orc.solve_batch(cart, xc, xref, dtype=np.float32, warm=warm, want_state=True, nthreads=os.cpu_count() or 1)
                if k >= 4:
                    bad["iter"] += int((ith_h[k - 4] != r.iter).sum())
                xc = orc.plant_step(cart, xc, r.u[:, 0, :], dtype=np.float32)
                warm = {q: r.state[q].copy() for q in ("d", "y", "g", "v", "z")}
                warm["y"][:] = 0
                warm["g"][:] = 0
            bad["plant_state"] = int((bt.get("x0")[:n] != xc).sum())
            chk_cl = {"instances": n, "mpc_steps": 4 + steps_cl, "bit_exact": all(v == 0 for v in bad.values()), "mismatching_elements": bad,
                      "compared": list(bad), "oracle_seconds": time.perf_counter() - t0c,
                      "how": "the oracle's closed loop (solve, its own plant step, duals reset) over all the steps for the first instances: iteration "
                             "counts of the timed steps and the final plant state (which every applied control went into)"}
        nx, nu, N = 4, 1, 10
        fused = s.stats()["launches"] == 2      # one persistent launch for all the steps + the last plant step (tmpc_kernel_small.cuh)
        state = 4 * (3 * nu * (N - 1) + 2 * nx * N)                    # d y z g v
        per_step_unfused = (4 * nx + state) + state + 4 * (nx * N + nu * (N - 1)) + 8 + 16 + 4 * nx    # in + out + x,u,iter,status,resid + new x0
        # fused: per INSTANCE and rollout x0, d, v, z in; workspace, x, u, iter, status, resid, x0 out; per step the iteration history
        per_rollout = (4 * nx + 4 * (2 * nu * (N - 1) + nx * N)) + state + 4 * (nx * N + nu * (N - 1)) + 8 + 16 + 2 * 4 * nx
        per_step = per_rollout / steps_cl + 4 if fused else per_step_unfused
        gbs = Bc * steps_cl * per_step / (ms_cl * 1e-3) / 1e9
        tf = iters_cl / world * FLOP_PER_ITER["c"] / (ms_cl * 1e-3) / 1e12
        bound = "hbm" if gbs / hbm_peak > tf / peak_tf else "fp32"
        out["config4_cartpole_closed_loop"] = {
            "workload": "the reference's cartpole closed loop (codegen_cartpole.cpp:75-122) for every instance, on the device: reset duals, "
                        "warm-started solve, plant step; x0 = 0.2 x the cold recipe's spread (|theta| <= 0.04 rad ...); %d MPC steps timed after 4 untimed ones" % steps_cl,
            "instances_total": T, "instances_per_gpu": Bc, "value": T * steps_cl / (ms_cl * 1e-3), "unit": "MPC steps/s (= solves/s)",
            "ms_per_mpc_step": ms_cl / steps_cl, "mean_iters_per_solve": iters_cl / (T * steps_cl), "iter_hist_0_to_7": hist,
            "fused": fused, "launches_per_rollout": s.stats()["launches"],
            "roofline": {"bound": bound, "achieved": gbs if bound == "hbm" else tf, "peak": hbm_peak if bound == "hbm" else peak_tf,
                         "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": gbs / hbm_peak if bound == "hbm" else tf / peak_tf,
                         "fp32": {"achieved_tflops": tf, "peak": peak_tf, "frac": tf / peak_tf, "algorithmic_flop_per_iteration": FLOP_PER_ITER["c"],
                                  "note": "FLOP of full iterations; the iteration that ends a step runs no backward sweep (admm.cpp:135-138)"},
                         "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak, "algorithmic_bytes_per_mpc_step": per_step,
                                 "algorithmic_bytes_per_mpc_step_one_launch_per_step": per_step_unfused, "peak_source": hbm_src,
                                 "bytes": "fused: the state stays in registers between the steps of a rollout, so per instance x0 + d, v, z in, workspace + "
                                          "x, u, iter, status, resid + x0 out ONCE per rollout (%d B / %d steps) + 4 B of iteration history per step; "
                                          "one launch per step (TMPC_ROLL=0): %d B per step" % (per_rollout, steps_cl, per_step_unfused)},
                         "traffic": None},
            "oracle_check": chk_cl}
        log("configs.config4_cartpole_closed_loop: %.3e MPC steps/s, frac %.3f (%s)%s" % (T * steps_cl / (ms_cl * 1e-3),
            out["config4_cartpole_closed_loop"]["roofline"]["frac"], bound, " fused" if fused else ""))
        bt.close(); s.close(); del ith
    except Exception as e:   # the entry is additional evidence: a failure here must not void the headline line
        out["config4_cartpole_closed_loop"] = {"error": repr(e)}
    torch.cuda.empty_cache()

    # ---- config 5: random 32/8/50 system, 262,144 instances in total: cold solve (untimed), x0 perturbed by 1 %, warm re-solve
    big = P.random_system()
    T = (1 << 18) // sc
    b0, b1 = pkg.sharding.shard_range(rank, world, total=T)
    x0, xref = W.random_system_batch(b0, b1)
    run = DeviceRun(torch, pkg, big, dev, local, args.policy, x0, xref, warm=True)
    run.step()                                    # solve #1, cold, leaves {d,y,g,v,z} in HBM
    torch.cuda.synchronize()
    saved = {k: v.clone() for k, v in run.warm.items()}
    x1 = W.perturb_x0(x0, b0)
    run.x0 = torch.from_numpy(x1).to(dev)

    def restore():
        for k, v in run.warm.items():
            v.copy_(saved[k])
    ms, st = run.timed(K, WU, barrier, before_each=restore)
    ms = allmax(ms)
    chk = None
    if rank == 0:
        n = min(CHECK_PREFIX, x0.shape[0])
        chk = oracle_check(big, x1, xref, outputs_np(run), n, warm_in={k: v[:n].cpu().numpy() for k, v in saved.items()})
    entry("config5_large_warm", "random 32/8/50 system (BASELINE configs[4]; seeded generator, problems.random_system): solve #1 cold (untimed), "
          "x0 perturbed by 1 %, solve #2 with {d,y,g,v,z} read from / written to HBM (timed; state restored between steps)", "l", run, T, ms, st,
          BYTES_PER_SOLVE["l_warm"], chk)
    run.close(); del run, saved; torch.cuda.empty_cache()
    return out


def multi_arm(torch, pkg, args, prob, config):
    """--impl multi (hand-run): ONE process, every visible GPU through tmpc_multi; prints one JSON line."""
    out = multi_measure(torch, pkg, args, prob, args.gpus if args.gpus > 1 else None)
    out["config"] = config
    emit(json.dumps(out))
    return 0


def multi_measure(torch, pkg, args, prob, devices):
    """ONE process, `devices` GPUs through tmpc_multi (one ctx + one worker thread per device): the product API's
    multi-GPU path for host callers (tiny_solve_batch / tmpc_multi_solve)."""
    capi = pkg.capi
    m = capi.Multi(prob, dtype=np.float32, policy=args.policy, devices=devices)
    G = m.device_count
    B = args.batch * G
    x0_np, xref_np = pkg.workloads.quadrotor_hover_batch(0, B, mult=args.mult)
    pin = lambda t: t.pin_memory()
    hx0, hxr = pin(torch.from_numpy(x0_np)), pin(torch.from_numpy(xref_np))
    hx = pin(torch.empty((B, prob.N, prob.nx), dtype=torch.float32))
    hu = pin(torch.empty((B, prob.N - 1, prob.nu), dtype=torch.float32))
    hu0 = pin(torch.empty((B, prob.nu), dtype=torch.float32))
    hit = pin(torch.empty(B, dtype=torch.int32))
    hst = pin(torch.empty(B, dtype=torch.int32))
    hrs = pin(torch.empty((B, 4), dtype=torch.float32))
    K = max(2, min(args.steps, 5))

    def timed(fn):
        for _ in range(2):
            fn()
        t = time.perf_counter()
        for _ in range(K):
            fn()
        return (time.perf_counter() - t) / K, m.stats()

    full_s, fs = timed(lambda: m.solve_raw(B, hx0, hxr, True, hx, hu, hit, hst, hrs))
    iters_full = int(hit.sum())
    u0_s, us = timed(lambda: m.solve_raw(B, hx0, hxr, True, None, None, hit, hst, None, u0=hu0))
    ok = iters_full == int(hit.sum()) and torch.equal(hu0, hu[:, 0, :])
    out = {"impl": "multi", "metric": "batched MPC solves/sec (quadrotor nx12 nu4 N10)", "n_gpus": G, "processes": 1, "unit": "solves/s",
           "instances": B, "steps": K,
           "e2e": {"value": B / full_s, "ms_per_step": 1e3 * full_s, "d2h_bytes_per_step": int(B * 648), "h2d_bytes_per_step": int(B * 48 + 480),
                   "slowest_kernel_ms": fs["kernel_ms"],
                   "u0_only": {"value": B / u0_s, "ms_per_step": 1e3 * u0_s, "d2h_bytes_per_step": int(B * 24), "slowest_kernel_ms": us["kernel_ms"]}},
           "u0_equals_u_col0": bool(ok), "iterations": iters_full,
           "how": "tmpc_multi_solve(TMPC_MEM_HOST) on pinned host arrays: contiguous instance ranges, one ctx + one host worker thread per device"}
    m.close()
    del hx, hu, hrs, hx0
    return out


if __name__ == "__main__":
    sys.exit(main())
