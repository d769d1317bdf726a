#!/usr/bin/env python3
"""Import the reference's example problem data into this repo's own `.mpcdata` text format.

Runs ONLY in the build container (needs /root/reference).  The outputs are committed under
accelerated-tinympc_b200/problem_data/ and are what the product, tests and bench read at run time.

Sources (row-major C arrays of `tinytype`, 7 decimals):
  examples/problem_data/quadrotor_{20,50,100}hz_params.hpp      A, B, Kinf, Pinf, Quu_inv, AmBKt, Q, R, rho
  examples/trajectory_data/quadrotor_20hz_y_axis_line.hpp        301 x 12 reference trajectory
  examples/codegen_cartpole.cpp:22-28                            cartpole A, B (column-major), Q, R, rho
`.mpcdata` stores every matrix COLUMN-MAJOR (the layout of the reference's structs, types.hpp:13-21).
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("TINYMPC_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "accelerated-tinympc_b200", "problem_data")


def c_arrays(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"tinytype\s+(\w+)\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}", txt):
        out[m.group(1)] = np.array([float(v) for v in re.split(r"[,\s]+", m.group(2).strip()) if v], dtype=np.float64)
    for m in re.finditer(r"tinytype\s+(\w+)\s*=\s*([-+0-9.eE]+)\s*;", txt):
        out[m.group(1)] = float(m.group(2))
    return out


def write_mpcdata(path, scalars, mats, comment):
    with open(path, "w") as f:
        f.write("# %s\n# matrices are column-major: value[r + c*rows]\n" % comment)
        for k, v in scalars.items():
            f.write("scalar %s %s\n" % (k, repr(v)))
        for k, m in mats.items():
            m = np.atleast_2d(m)
            f.write("matrix %s %d %d\n" % (k, m.shape[0], m.shape[1]))
            f.write(" ".join(repr(float(x)) for x in m.flatten(order="F")) + "\n")


def quadrotor(hz):
    a = c_arrays(os.path.join(REF, "examples/problem_data/quadrotor_%dhz_params.hpp" % hz))
    nx, nu = 12, 4
    rm = lambda name, r, c: a[name].reshape(r, c)  # headers are row-major (quadrotor_hovering.cpp:33-41)
    mats = {
        "Adyn": rm("Adyn_data", nx, nx), "Bdyn": rm("Bdyn_data", nx, nu),
        "Kinf": rm("Kinf_data", nu, nx), "Pinf": rm("Pinf_data", nx, nx),
        "Quu_inv": rm("Quu_inv_data", nu, nu), "AmBKt": rm("AmBKt_data", nx, nx),
        "Q": a["Q_data"].reshape(nx, 1), "R": a["R_data"].reshape(nu, 1),
    }
    write_mpcdata(os.path.join(OUT, "quadrotor_%dhz.mpcdata" % hz), {"nx": nx, "nu": nu, "rho": a["rho_value"]}, mats,
                  "quadrotor %d Hz (from examples/problem_data/quadrotor_%dhz_params.hpp)" % (hz, hz))


def trajectory():
    a = c_arrays(os.path.join(REF, "examples/trajectory_data/quadrotor_20hz_y_axis_line.hpp"))
    X = a["Xref_data"].reshape(301, 12).T  # -> nx x NTOTAL, as quadrotor_tracking.cpp:84
    write_mpcdata(os.path.join(OUT, "quadrotor_20hz_y_axis_line.mpcdata"), {"nx": 12, "ntotal": 301}, {"Xref_total": X},
                  "20 Hz y-axis line trajectory (from examples/trajectory_data/quadrotor_20hz_y_axis_line.hpp)")


def cartpole():
    txt = open(os.path.join(REF, "examples/codegen_cartpole.cpp")).read()
    def arr(name):
        m = re.search(r"tinytype\s+%s\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}" % name, txt)
        return np.array([float(v) for v in re.split(r"[,\s]+", m.group(1).strip()) if v])
    rho = float(re.search(r"tinytype\s+rho_value\s*=\s*([-+0-9.eE]+)", txt).group(1))
    n, m = 4, 1
    mats = {"Adyn": arr("Adyn_data").reshape(n, n, order="F"), "Bdyn": arr("Bdyn_data").reshape(n, m, order="F"),
            "Q": arr("Q_data").reshape(n, 1), "R": arr("R_data").reshape(m, 1)}
    write_mpcdata(os.path.join(OUT, "cartpole.mpcdata"), {"nx": n, "nu": m, "rho": rho}, mats,
                  "cartpole model, no cache (from examples/codegen_cartpole.cpp:22-28; cache comes from tiny_precompute)")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not found at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    for hz in (20, 50, 100):
        quadrotor(hz)
    trajectory()
    cartpole()
    print("wrote", sorted(os.listdir(OUT)))
