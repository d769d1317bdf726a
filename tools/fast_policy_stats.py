"""FAST policy (FMA-contracted, sequential accumulation) against the bit-exact oracle: iteration-count mismatch rate and |delta iter| histogram, to
set beside the reference's own flag-to-flag spread (float -O3 SSE2 vs -mavx2 -mfma: 2.3 % mismatches, max |delta| 11)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import OracleLib  # noqa: E402  (checker)

pkg = load_package()
prob = pkg.problems.quadrotor(20)
B = 100_000
for mult in (0.25, 1.0):
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, B, mult=mult)
    ref = OracleLib().solve_batch(prob, x0, xref, dtype=np.float32, nthreads=os.cpu_count() or 1)
    out = pkg.capi.Solver(prob, dtype=np.float32, policy="fast").solve(x0, xref)
    d = np.abs(out["iter"].astype(np.int64) - ref.iter)
    same = d == 0
    scale = np.maximum(np.abs(ref.x[same]).max(), 1.0)
    print("mult %.2f: %d instances, iteration-count mismatches %.3f %%, max |delta iter| %d, histogram of |delta| 1..8+: %s; on equal-iteration instances max rel "
          "|dx| %.2e, max |du| %.2e" % (mult, B, 100.0 * (~same).mean(), d.max(), np.bincount(np.minimum(d, 8), minlength=9)[1:].tolist(),
                                       np.abs(out["x"][same] - ref.x[same]).max() / scale, np.abs(out["u"][same] - ref.u[same]).max()), flush=True)
