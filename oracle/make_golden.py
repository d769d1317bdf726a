#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: generate the golden fixtures under tests/golden/ from the UNMODIFIED reference
(oracle/_ref/libref_*.so, built by oracle/Makefile from /root/reference).  Runs only in the build
container; the fixtures are committed and are what the GPU box checks against.

Fixtures (npz, small):
  hover_closed_loop_{f32,f64}.npz    examples/quadrotor_hovering.cpp:83-114 replayed: 70 MPC steps, warm start
                                     carried exactly like the example (y, g reset each step; d, v, z kept)
  tracking_closed_loop_{f32,f64}.npz examples/quadrotor_tracking.cpp:84-118: 290 steps on the sliding Xref window
  cartpole_closed_loop_f32.npz       examples/codegen_cartpole.cpp:75-122: 300 steps, max_iter 150
  batch_{q,c,l}_{f32,f64}.npz        seeded batches: per-instance iter/status/resid + x,u of a prefix
  steps_{q,c,l}_{f32,f64}.npz        the six step functions (admm.hpp:13-18) on random workspaces
The plant simulation between MPC steps (x+ = A x + B u0) is done here in float64 and the resulting x0 of every
step is STORED, so replays feed bit-identical inputs on any machine.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import RefLib, ws_size  # noqa: E402

pkg = load_package()
P, W = pkg.problems, pkg.workloads
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
DT = {"f32": np.float32, "f64": np.float64}


def closed_loop(ref, prob, x0, xref_of_step, steps, dtype):
    nx, nu, N = prob.nx, prob.nu, prob.N
    warm = {k: np.zeros((1, N - 1, nu) if k in "dyz" else (1, N, nx), dtype) for k in ("d", "y", "g", "v", "z")}
    rec = {k: [] for k in ("x0", "xref", "iter", "status", "u0", "resid", "x_last")}
    A, B = prob.Adyn, prob.Bdyn
    x0 = np.asarray(x0, np.float64)
    for k in range(steps):
        xr = np.asarray(xref_of_step(k), dtype)
        x0c = x0.astype(dtype)
        warm["y"][:] = 0
        warm["g"][:] = 0
        r = ref.solve_batch(prob, x0c[None, :], xr, warm=warm, want_state=True)
        warm = {kk: r.state[kk].copy() for kk in ("d", "y", "g", "v", "z")}
        rec["x0"].append(x0c); rec["xref"].append(xr); rec["iter"].append(r.iter[0]); rec["status"].append(r.status[0])
        rec["u0"].append(r.u[0, 0].copy()); rec["resid"].append(r.resid[0].copy()); rec["x_last"].append(r.x[0, -1].copy())
        x0 = A @ x0c.astype(np.float64) + B @ r.u[0, 0].astype(np.float64)
    return {k: np.array(v) for k, v in rec.items()}


def gen_closed_loops():
    q = P.quadrotor(20)
    hover = np.tile(W.QUAD_HOVER[None, :], (q.N, 1))
    table = P.quadrotor_trajectory()  # (12, 301)
    for tag, dt in DT.items():
        ref = RefLib("q_" + tag)
        x0 = np.array([0, 1, 0, 0.2, 0, 0, 0.1, 0, 0, 0, 0, 0], np.float64)          # quadrotor_hovering.cpp:88
        rec = closed_loop(ref, q, x0, lambda k: hover, 70, dt)
        np.savez_compressed(os.path.join(OUT, "hover_closed_loop_%s.npz" % tag), **rec)
        print("hover", tag, "iters:", " ".join(str(i) for i in rec["iter"]))
        rec = closed_loop(ref, q, table[:, 0], lambda k: table[:, k:k + q.N].T, 290, dt)  # tracking.cpp:84-101
        np.savez_compressed(os.path.join(OUT, "tracking_closed_loop_%s.npz" % tag), **rec)
        print("tracking", tag, "iter hist:", np.bincount(rec["iter"]).nonzero()[0], np.bincount(rec["iter"])[np.bincount(rec["iter"]).nonzero()[0]])
    c = P.cartpole(max_iter=150)                                                        # cartpole.cpp:81
    ref = RefLib("c_f32")
    rec = closed_loop(ref, c, np.array([0.0, 0, 0.1, 0]), lambda k: np.zeros((c.N, 4)), 300, np.float32)
    np.savez_compressed(os.path.join(OUT, "cartpole_closed_loop_f32.npz"), **rec)
    print("cartpole f32 iter hist:", np.bincount(rec["iter"]))


def rollout(ref, prob, x0s, table, starts, steps, dtype):
    """The examples' closed loop, verbatim, for a few instances: measurement, reference window from a table
    (quadrotor_tracking.cpp:101; a one-window table = hovering), duals reset, tiny_solve, plant step on the
    reference's own Eigen expression (ref_plant_step).  Everything in `dtype`, nothing computed in numpy."""
    nx, nu, N = prob.nx, prob.nu, prob.N
    B = len(x0s)
    rec = {k: [] for k in ("x0", "iter", "status", "u0", "err")}
    for b in range(B):
        warm = {k: np.zeros((1, N - 1, nu) if k in "dyz" else (1, N, nx), dtype) for k in ("d", "y", "g", "v", "z")}
        x0 = np.asarray(x0s[b], dtype)
        xs, its, sts, u0s, errs = [x0.copy()], [], [], [], []
        for k in range(steps):
            w0 = min(starts[b] + k, table.shape[0] - N)
            xr = np.ascontiguousarray(table[w0:w0 + N], dtype)
            warm["y"][:] = 0
            warm["g"][:] = 0
            r = ref.solve_batch(prob, x0[None, :], xr, warm=warm, want_state=True)
            warm = {kk: r.state[kk].copy() for kk in ("d", "y", "g", "v", "z")}
            x1, err = ref.plant_step(prob, x0, r.u[0], xref=xr)
            its.append(r.iter[0]); sts.append(r.status[0]); u0s.append(r.u[0, 0].copy()); errs.append(err)
            x0 = x1
            xs.append(x0.copy())
        rec["x0"].append(np.array(xs)); rec["iter"].append(np.array(its)); rec["status"].append(np.array(sts))
        rec["u0"].append(np.array(u0s)); rec["err"].append(np.array(errs))
    out = {k: np.array(v) for k, v in rec.items()}          # [B][steps(+1)][...]
    out["starts"] = np.asarray(starts, np.int32)
    out["x0_init"] = np.asarray(x0s, dtype)
    return out


def gen_rollouts():
    """Closed loops INCLUDING the reference's plant step: what tmpc_batch_rollout must reproduce on the device."""
    q = P.quadrotor(20)
    hover_tab = np.tile(W.QUAD_HOVER[None, :], (q.N, 1))           # one window: Xref never moves
    table = P.quadrotor_trajectory().T                              # [301][12]
    for tag, dt in DT.items():
        ref = RefLib("q_" + tag)
        x0s = np.concatenate([np.array([[0, 1, 0, 0.2, 0, 0, 0.1, 0, 0, 0, 0, 0]], np.float64),   # hovering.cpp:88
                              W.quadrotor_hover_batch(0, 7, mult=0.5)[0].astype(np.float64)])
        rec = rollout(ref, q, x0s, hover_tab, [0] * 8, 70, dt)
        np.savez_compressed(os.path.join(OUT, "rollout_hover_%s.npz" % tag), **rec)
        print("rollout hover", tag, "instance 0 iters:", " ".join(str(i) for i in rec["iter"][0][:16]), "err0 %.4f" % rec["err"][0][0])
        starts = [0, 17, 100, 250]
        x0s = np.array([table[s0] for s0 in starts], np.float64)
        x0s[1:] += 0.05 * W.QUAD_SCALE[None, :] * W._noise(4321, 0, 3, 12)
        rec = rollout(ref, q, x0s, table, starts, 60, dt)
        np.savez_compressed(os.path.join(OUT, "rollout_tracking_%s.npz" % tag), **rec)
        print("rollout tracking", tag, "iters[0][:8]:", rec["iter"][0][:8])
    c = P.cartpole(max_iter=150)
    ref = RefLib("c_f32")
    x0s = np.concatenate([np.array([[0.0, 0, 0.1, 0]]), W.cartpole_batch(0, 5)[0].astype(np.float64)])
    rec = rollout(ref, c, x0s, np.zeros((c.N, 4)), [0] * 6, 80, np.float32)
    np.savez_compressed(os.path.join(OUT, "rollout_cartpole_f32.npz"), **rec)
    print("rollout cartpole f32 iter hist:", np.bincount(rec["iter"].reshape(-1)))


def gen_codegen():
    """What the reference's tiny_codegen (codegen.cpp:218) emits for the cartpole model of examples/codegen_cartpole.cpp
    and for the 20 Hz quadrotor model: src/tiny_data_workspace.cpp and tinympc/glob_opts.hpp, timestamp line removed."""
    import shutil
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "codegen_ref")
    cs, cm = P.cartpole_model()
    q = P.quadrotor(20)
    jobs = {"cartpole": (4, 1, 10, float(cs["rho"]), cm["Adyn"], cm["Bdyn"], cm["Q"].reshape(-1), cm["R"].reshape(-1), 5.0, 5.0, 100),
            "quadrotor": (12, 4, 10, q.rho, q.Adyn, q.Bdyn, q.Q, q.R, 5.0, 0.5, 100)}
    for name, (nx, nu, N, rho, A, B, Q, R, xb, ub, max_iter) in jobs.items():
        tmp = tempfile.mkdtemp(prefix="tmpc_codegen_")
        with open(os.path.join(tmp, "problem.txt"), "w") as f:
            f.write("%d %d %d %.17g 1e-3 1e-3 %d 1\n" % (nx, nu, N, rho, max_iter))
            for arr in (np.asarray(A).flatten(order="F"), np.asarray(B).flatten(order="F"), Q, R, np.full(nx * N, -xb), np.full(nx * N, xb),
                        np.full(nu * (N - 1), -ub), np.full(nu * (N - 1), ub)):
                f.write(" ".join("%.17g" % v for v in np.asarray(arr, np.float64).reshape(-1)) + "\n")
        out = os.path.join(tmp, "generated")
        subprocess.run([exe, os.path.join(tmp, "problem.txt"), "/root/reference", out], check=True, stdout=subprocess.DEVNULL)
        for src, dst in (("src/tiny_data_workspace.cpp", "codegen_%s_tiny_data_workspace.cpp.txt" % name),
                         ("tinympc/glob_opts.hpp", "codegen_%s_glob_opts.hpp.txt" % name)):
            lines = [ln for ln in open(os.path.join(out, src)).read().split("\n") if "autogenerated by TinyMPC on" not in ln]
            open(os.path.join(OUT, dst), "w").write("\n".join(lines))
        shutil.rmtree(tmp)
        print("codegen golden:", name)


def gen_batches():
    q, c, l = P.quadrotor(20), P.cartpole(), P.random_system()
    rng = np.random.default_rng(7)
    cases = []
    for mult in (0.1, 0.25, 1.0):
        x0, xr = W.quadrotor_hover_batch(0, 2000, mult=mult)
        cases.append(("q", "hover_m%s" % mult, q, x0, xr))
    x0, xr = W.quadrotor_tracking_batch(0, 1160)
    cases.append(("q", "tracking", q, x0, xr))
    x0, xr = W.cartpole_batch(0, 4000)
    cases.append(("c", "cartpole", c, x0, xr))
    x0 = rng.uniform(-1, 1, (96, 32)).astype(np.float32)
    cases.append(("l", "random32", l, x0, np.zeros((50, 32), np.float32)))
    for tag, dt in DT.items():
        out = {}
        for shape, name, prob, x0, xr in cases:
            ref = RefLib("%s_%s" % (shape, tag))
            r = ref.solve_batch(prob, x0, xr, nthreads=8)
            keep = 64
            out.update({name + "_iter": r.iter, name + "_status": r.status, name + "_resid": r.resid,
                        name + "_x": r.x[:keep], name + "_u": r.u[:keep],
                        name + "_xsum": r.x.astype(np.float64).sum(axis=0), name + "_usum": r.u.astype(np.float64).sum(axis=0)})
            print("batch", name, tag, "mean iters %.3f solved %d/%d" % (r.iter.mean(), (r.status == 1).sum(), len(r.iter)))
        np.savez_compressed(os.path.join(OUT, "batch_%s.npz" % tag), **out)


def gen_steps():
    probs = {"q": P.quadrotor(20), "c": P.cartpole(), "l": P.random_system()}
    rng = np.random.default_rng(11)
    for shape, prob in probs.items():
        n = ws_size(prob.nx, prob.nu, prob.N)
        wss = rng.uniform(-1, 1, (2 if shape == "l" else 6, n))
        for tag, dt in DT.items():
            ref = RefLib("%s_%s" % (shape, tag))
            out = {"ws_in": wss.astype(dt)}
            for which in range(6):
                res = [ref.step(prob, which, w.astype(dt), it=1) for w in wss]
                out["out%d" % which] = np.array([r[1] for r in res])
                out["rc%d" % which] = np.array([r[0] for r in res], np.int32)
            np.savez_compressed(os.path.join(OUT, "steps_%s_%s.npz" % (shape, tag)), **out)
    print("steps done")


if __name__ == "__main__":
    which = sys.argv[1:] or ["closed_loops", "batches", "steps", "rollouts", "codegen"]
    if "closed_loops" in which:
        gen_closed_loops()
    if "batches" in which:
        gen_batches()
    if "steps" in which:
        gen_steps()
    if "rollouts" in which:
        gen_rollouts()
    if "codegen" in which:
        gen_codegen()
    print("fixtures:", sorted(os.listdir(OUT)))
