#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Pins select_orders() (oracle/tinympc_oracle.c) for shapes beyond the three shipped ones:
for every (nx, nu, N) listed it builds the UNMODIFIED reference for that shape (`make -C oracle refshape ...`,
needs /root/reference), then compares the plain-C oracle with it bit for bit on seeded full solves (whole
workspace) and on each of the six step functions, f32 and f64.  A step-function mismatch names the product whose
evaluation order is wrong.

  python oracle/pin_shapes.py                 # the default shape list
  python oracle/pin_shapes.py 6,3,20 9,2,7    # chosen shapes
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import OracleLib, RefLib, ws_size  # noqa: E402

DEFAULT = [(2, 1, 5), (3, 1, 8), (4, 2, 10), (5, 2, 6), (6, 3, 20), (7, 3, 9), (8, 2, 15), (8, 4, 12), (8, 8, 6),
           (9, 4, 7), (10, 5, 10), (12, 6, 8), (13, 7, 6), (16, 4, 20), (16, 8, 25), (17, 9, 5), (20, 8, 12),
           (24, 6, 10), (24, 12, 8), (31, 7, 6), (40, 10, 6), (6, 6, 6), (3, 3, 4), (16, 16, 5), (8, 1, 10),
           (12, 1, 10), (5, 1, 6), (12, 8, 5), (12, 9, 5)]
STEP_NAMES = ["forward_pass", "update_slack", "update_dual", "update_linear_cost", "termination_condition",
              "backward_pass_grad"]
DT = {"f32": np.float32, "f64": np.float64}


def build(shape, sc):
    nx, nu, N = shape
    cfg = "g%dx%dx%d_%s" % (nx, nu, N, sc)
    if not RefLib.available(cfg):
        subprocess.check_call(["make", "-s", "-C", HERE, "refshape", "NX=%d" % nx, "NU=%d" % nu, "NH=%d" % N,
                               "SC=" + sc], stdout=subprocess.DEVNULL)
    return cfg


def cleanup(cfg):
    """Generic-shape reference builds are scratch: they would otherwise travel to the GPU box with every snapshot."""
    import shutil
    try:
        os.remove(os.path.join(HERE, "_ref", "libref_%s.so" % cfg))
    except OSError:
        pass
    shutil.rmtree(os.path.join(HERE, "_ref", "stage", cfg), ignore_errors=True)


def check(pkg, ora, shape, sc, keep=False):
    nx, nu, N = shape
    cfg = build(shape, sc)
    ref = RefLib(cfg)
    if not keep:
        cleanup(cfg)   # the library stays mapped in this process
    prob = pkg.problems.random_system(nx, nu, N, seed=100 + nx * 7 + nu)
    rng = np.random.default_rng(nx * 1000 + nu * 10 + N)
    bad = []
    n = ws_size(nx, nu, N)
    for trial in range(6):
        ws = rng.uniform(-1, 1, n).astype(DT[sc])
        for which in range(6):
            rc_r, out_r = ref.step(prob, which, ws, it=1)
            rc_o, out_o = ora.step(prob, which, ws, it=1, dtype=DT[sc])
            if rc_r != rc_o or not np.array_equal(out_r, out_o):
                if STEP_NAMES[which] not in bad:
                    bad.append(STEP_NAMES[which])
    # the examples' plant step x1 = Adyn * x0 + Bdyn * u.col(0) (quadrotor_hovering.cpp:108)
    xp = rng.uniform(-2, 2, (16, nx)).astype(DT[sc])
    up = rng.uniform(-1, 1, (16, N - 1, nu)).astype(DT[sc])
    exp = np.stack([ref.plant_step(prob, xp[b], up[b]) for b in range(16)])
    if not np.array_equal(ora.plant_step(prob, xp, up[:, 0], dtype=DT[sc]), exp):
        bad.append("plant_step")
    x0 = rng.uniform(-3, 3, (200, nx)).astype(np.float32)
    xref = rng.uniform(-0.5, 0.5, (N, nx)).astype(np.float32)
    r = ref.solve_batch(prob, x0, xref, want_state=True, nthreads=2)
    o = ora.solve_batch(prob, x0, xref, dtype=DT[sc], want_state=True, nthreads=2)
    for name in ("iter", "status", "x", "u", "resid"):
        if not np.array_equal(getattr(o, name), getattr(r, name)):
            bad.append("solve." + name)
    for k in r.state:
        if not np.array_equal(o.state[k], r.state[k]):
            bad.append("ws." + k)
    return bad, float(r.iter.mean()), float((r.status == 1).mean())


def main():
    shapes = [tuple(int(t) for t in a.split(",")) for a in sys.argv[1:]] or DEFAULT
    pkg = load_package()
    ora = OracleLib()
    jobs = [(s, sc) for s in shapes for sc in ("f32", "f64")]
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(lambda j: build(*j), jobs))
    nbad = 0
    for s, sc in jobs:
        bad, mean_it, solved = check(pkg, ora, s, sc)
        nbad += bool(bad)
        print("%-12s %s  mean_iter %5.1f solved %4.2f  %s" % ("%d/%d/%d" % s, sc, mean_it, solved,
                                                               "OK" if not bad else "MISMATCH: " + ", ".join(bad)))
    print("%d of %d shape/scalar combinations differ" % (nbad, len(jobs)))
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
