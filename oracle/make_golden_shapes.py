#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: golden fixtures for shapes BEYOND the three BASELINE ones, generated from the UNMODIFIED
reference compiled for each shape (`make -C oracle refshape ...`, needs /root/reference; build container only).

  tests/golden/shapes_{f32,f64}.npz   per shape "nx_nu_N": the model (so no linear algebra has to reproduce on another
                                      machine), x0 [B][nx], Xref [N][nx], and the reference's iter / status / resid / x / u,
                                      plus a warm-started second solve (x0 * 1.01 from the first solve's d y g v z).
The shapes cover every branch of the evaluation-order dispatch (tmpc_orders_rt.hpp): nx, nu multiples of the SSE packet or
not, unrolled or address-peeled assignments, single input, GEMV threshold, K beyond the unrolling limit.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

SHAPES = [(2, 2, 3), (2, 1, 5), (3, 1, 8), (5, 2, 6), (6, 3, 20), (7, 3, 9), (8, 2, 15), (9, 4, 7), (10, 5, 10), (12, 6, 8), (13, 7, 6),
          (16, 8, 25), (17, 9, 5), (24, 6, 10), (31, 7, 6), (40, 10, 6), (5, 1, 6), (12, 9, 5), (60, 1, 4), (57, 3, 4), (64, 16, 3)]
DT = {"f32": np.float32, "f64": np.float64}
B = 32
MODEL_KEYS = ("Adyn", "Bdyn", "Q", "Kinf", "Pinf", "Quu_inv", "AmBKt", "x_min", "x_max", "u_min", "u_max")


def problem(pkg, shape):
    nx, nu, N = shape
    if shape == (2, 2, 3):
        return pkg.problems.codegen_random()   # the reference's own examples/codegen_random.cpp data
    return pkg.problems.random_system(nx, nu, N, seed=100 + nx * 7 + nu)


def main():
    pkg = load_package()
    for tag, dt in DT.items():
        rec = {}
        for shape in SHAPES:
            nx, nu, N = shape
            cfg = "g%dx%dx%d_%s" % (nx, nu, N, tag)
            if not RefLib.available(cfg):
                subprocess.check_call(["make", "-s", "-C", HERE, "refshape", "NX=%d" % nx, "NU=%d" % nu, "NH=%d" % N, "SC=" + tag],
                                      stdout=subprocess.DEVNULL)
            ref = RefLib(cfg)
            from oracle.pin_shapes import cleanup
            cleanup(cfg)   # scratch build; stays mapped in this process
            prob = problem(pkg, shape)
            rng = np.random.default_rng(nx * 1000 + nu * 10 + N)
            x0 = rng.uniform(-3, 3, (B, nx)).astype(np.float32)
            x0[: B // 2] *= np.float32(0.1)   # half of the batch close enough to converge early
            xref = rng.uniform(-0.5, 0.5, (N, nx)).astype(np.float32)
            r1 = ref.solve_batch(prob, x0, xref, want_state=True, nthreads=2)
            warm = {k: r1.state[k] for k in ("d", "y", "g", "v", "z")}
            r2 = ref.solve_batch(prob, (x0 * 1.01).astype(np.float32), xref, warm=warm, want_state=True, nthreads=2)
            key = "%d_%d_%d/" % shape
            for k in MODEL_KEYS:
                rec[key + k] = np.asarray(getattr(prob, k), np.float64)
            rec[key + "rho"] = np.float64(prob.rho)
            rec[key + "x0"] = x0
            rec[key + "xref"] = xref
            for name in ("iter", "status", "resid", "x", "u"):
                rec[key + name] = getattr(r1, name)
                rec[key + "w_" + name] = getattr(r2, name)
            for k in ("d", "y", "g", "v", "z"):
                rec[key + "w_state_" + k] = r2.state[k]
            print(tag, shape, "mean iter %.1f / warm %.1f, solved %.2f" % (r1.iter.mean(), r2.iter.mean(), (r1.status == 1).mean()))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "shapes_%s.npz" % tag), **rec)


if __name__ == "__main__":
    main()
