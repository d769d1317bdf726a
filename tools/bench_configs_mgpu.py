"""BASELINE.json configs 3-5 at the GPU counts they name, one rank per GPU (launch with torch.distributed.run like
bench.py).  The batch shards by contiguous index ranges (sharding.py), no data-path collective; per config: 2 warm-up
steps, K timed steps bracketed by barrier + synchronize, CUDA events on the launch stream, MAX over ranks; rank 0
prints one JSON line per config.

  config 3  quadrotor_tracking, per-instance Xref windows, 4,194,304 instances TOTAL (strong scaling over the ranks)
  config 4  codegen_cartpole 4/1/10, 16,777,216 instances total
  config 5  random 32/8/50, 262,144 instances total, cold solve (untimed) then x0 perturbed 1 % and warm re-solve (timed)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_configs_mgpu.py [steps]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
pkg = load_package()
capi = pkg.capi
FLOP = {"q": 11918, "c": 1771, "l": 344058}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak_tf = 148 * 128 * 2 * 1.965e9 / 1e12

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(name, shape, prob, total, gen, shared, warm_sequence=False, policy="parity"):
        b0, b1 = pkg.sharding.shard_range(rank, world, total=total)
        B = b1 - b0
        x0, xref = gen(b0, b1)
        s = capi.Solver(prob, dtype=np.float32, policy=policy, device=local)
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        x0d, xrd = f(x0), f(xref)
        x = torch.empty((B, prob.N, prob.nx), device=dev); u = torch.empty((B, prob.N - 1, prob.nu), device=dev)
        it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
        rs = torch.empty((B, 4), device=dev)
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        warm = saved = None
        if warm_sequence:
            warm = {k: torch.zeros((B, prob.N - 1, prob.nu) if k in "dyz" else (B, prob.N, prob.nx), device=dev) for k in ("d", "y", "g", "v", "z")}
            s.solve_raw(B, x0d, xrd, shared, capi.TMPC_MEM_DEVICE, x, u, it, st, rs, warm=warm, stream=stream.cuda_stream)   # cold leg, untimed
            torch.cuda.synchronize()
            saved = {k: v.clone() for k, v in warm.items()}
            x0d = f(pkg.workloads.perturb_x0(x0, b0))

        def step():
            if saved is not None:
                for k in warm:
                    warm[k].copy_(saved[k], non_blocking=True)     # restoring the carried state is not part of the solve: see kernel_ms
            s.solve_raw(B, x0d, xrd, shared, capi.TMPC_MEM_DEVICE, x, u, it, st, rs, warm=warm, stream=stream.cuda_stream)

        for _ in range(3):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kms = 0.0
        for _ in range(steps):
            step()
            if saved is not None:
                torch.cuda.synchronize()
                kms += s.stats()["kernel_ms"]
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) if saved is None else kms          # warm sequence: sum of the solve kernels only
        stt = s.stats()
        vec = pkg.sharding.local_stats(it.cpu().numpy(), st.cpu().numpy(), prob.max_iter)
        vec, tmax = pkg.sharding.gather_stats(vec, [ms], dist if world > 1 else None, dev)
        ms = float(tmax[0])
        iters, solved, inst = float(vec[0]), float(vec[1]), float(vec[2])
        if rank == 0:
            tf = iters * steps * FLOP[shape] / (ms * 1e-3) / 1e12
            rec = {"name": name, "n_gpus": world, "instances_total": int(inst), "steps": steps, "ms_per_step": ms / steps,
                   "solves_per_s": inst * steps / (ms * 1e-3), "iters_per_s": iters * steps / (ms * 1e-3), "mean_iters": iters / inst,
                   "solved_frac": solved / inst, "policy": policy, "algorithmic_tflops": tf, "fp32_roofline_frac_per_gpu": tf / world / peak_tf,
                   "timing": "CUDA events over the timed steps, max over ranks" if saved is None else "sum of solve-kernel times, max over ranks",
                   "pattern": stt.get("pattern", 0)}
            os.write(REAL_STDOUT, (json.dumps(rec) + "\n").encode())
        s.close()
        del x, u, it, st, rs, x0d, xrd, warm, saved
        torch.cuda.empty_cache()

    W = pkg.workloads
    quad, cart, big = pkg.problems.quadrotor(20), pkg.problems.cartpole(), pkg.problems.random_system()
    run("config3_tracking_4M", "q", quad, 4 * (1 << 20), lambda a, b: W.quadrotor_tracking_batch(a, b), False)
    run("config4_cartpole_16M", "c", cart, 16 * (1 << 20), lambda a, b: W.cartpole_batch(a, b), True)
    run("config5_large_262144_warm", "l", big, 262144, lambda a, b: W.random_system_batch(a, b), True, warm_sequence=True)
    run("config2_hover_1M_per_gpu", "q", quad, world * (1 << 20), lambda a, b: W.quadrotor_hover_batch(a, b, mult=0.25), True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
