"""Problem definitions for the batched TinyMPC path: model + cache + bounds + settings.

Mirrors what the reference keeps in TinyCache / TinyWorkspace params / TinySettings
(/root/reference/src/tinympc/types.hpp:26-97) and how its examples fill them
(examples/quadrotor_hovering.cpp:33-78).  Everything is COLUMN-MAJOR on the wire, i.e. a numpy array
of shape (rows, cols) must be passed with order="F"; trajectories are stored as [stage][dim]
C-contiguous arrays (= the column-major nx x N matrix of the reference, tiny_wrapper.cpp:27).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "problem_data")


def read_mpcdata(path):
    """Parse a .mpcdata file -> (scalars: dict, matrices: dict of float64 (rows, cols) arrays)."""
    scalars, mats = {}, {}
    with open(path) as f:
        lines = [ln.strip() for ln in f if ln.strip() and not ln.startswith("#")]
    i = 0
    while i < len(lines):
        tok = lines[i].split()
        if tok[0] == "scalar":
            v = float(tok[2])
            scalars[tok[1]] = int(v) if v.is_integer() and tok[1] != "rho" else v
            i += 1
        elif tok[0] == "matrix":
            r, c = int(tok[2]), int(tok[3])
            vals = np.array([float(x) for x in lines[i + 1].split()], dtype=np.float64)
            if vals.size != r * c:
                raise ValueError("%s: matrix %s has %d values, expected %d" % (path, tok[1], vals.size, r * c))
            mats[tok[1]] = vals.reshape(r, c, order="F")
            i += 2
        else:
            raise ValueError("%s: bad line %r" % (path, lines[i]))
    return scalars, mats


@dataclass
class Problem:
    """One MPC problem family shared by every instance of a batch (shared cache)."""
    nx: int
    nu: int
    N: int
    rho: float
    Adyn: np.ndarray      # (nx, nx)
    Bdyn: np.ndarray      # (nx, nu)
    Q: np.ndarray         # (nx,)   used as given by update_linear_cost (admm.cpp:81)
    Kinf: np.ndarray      # (nu, nx)
    Pinf: np.ndarray      # (nx, nx)
    Quu_inv: np.ndarray   # (nu, nu)
    AmBKt: np.ndarray     # (nx, nx)
    x_min: np.ndarray | None = None   # (N, nx)   [stage][dim]
    x_max: np.ndarray | None = None
    u_min: np.ndarray | None = None   # (N-1, nu)
    u_max: np.ndarray | None = None
    abs_pri_tol: float = 1e-3
    abs_dua_tol: float = 1e-3
    max_iter: int = 100
    check_termination: int = 1
    en_state_bound: int = 1
    en_input_bound: int = 1
    R: np.ndarray | None = None
    name: str = ""
    extra: dict = field(default_factory=dict)

    def with_bounds(self, xlo, xhi, ulo, uhi):
        self.x_min = np.full((self.N, self.nx), float(xlo))
        self.x_max = np.full((self.N, self.nx), float(xhi))
        self.u_min = np.full((self.N - 1, self.nu), float(ulo))
        self.u_max = np.full((self.N - 1, self.nu), float(uhi))
        return self

    def cast(self, dtype):
        """Arrays as flat column-major buffers of `dtype` (what the C ABIs take)."""
        dt = np.dtype(dtype)
        f = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).flatten(order="F").astype(dt))
        c = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(dt).reshape(-1))
        return {
            "Kinf": f(self.Kinf), "Pinf": f(self.Pinf), "Quu_inv": f(self.Quu_inv), "AmBKt": f(self.AmBKt),
            "Adyn": f(self.Adyn), "Bdyn": f(self.Bdyn), "Q": c(self.Q),
            "x_min": c(self.x_min), "x_max": c(self.x_max), "u_min": c(self.u_min), "u_max": c(self.u_max),
        }


def quadrotor(hz: int = 20, N: int = 10) -> Problem:
    """Reference quadrotor example: shipped cache, rho = 5, |u| <= 0.5, |x| <= 5, tol 1e-3, max_iter 100
    (examples/quadrotor_hovering.cpp:33-47, 73-78)."""
    s, m = read_mpcdata(os.path.join(DATA_DIR, "quadrotor_%dhz.mpcdata" % hz))
    p = Problem(nx=s["nx"], nu=s["nu"], N=N, rho=float(s["rho"]), Adyn=m["Adyn"], Bdyn=m["Bdyn"],
                Q=m["Q"].reshape(-1), R=m["R"].reshape(-1), Kinf=m["Kinf"], Pinf=m["Pinf"],
                Quu_inv=m["Quu_inv"], AmBKt=m["AmBKt"], name="quadrotor_%dhz" % hz)
    return p.with_bounds(-5.0, 5.0, -0.5, 0.5)


def quadrotor_trajectory() -> np.ndarray:
    """(12, 301) reference table of examples/quadrotor_tracking.cpp:84."""
    _, m = read_mpcdata(os.path.join(DATA_DIR, "quadrotor_20hz_y_axis_line.mpcdata"))
    return m["Xref_total"]


def cartpole_model():
    """Cartpole A, B, Q, R, rho (examples/codegen_cartpole.cpp:22-28); the cache comes from precompute."""
    s, m = read_mpcdata(os.path.join(DATA_DIR, "cartpole.mpcdata"))
    return s, m


def precompute_cache(A, B, Q, R, rho, max_riccati_iter: int = 1000, tol: float = 1e-5):
    """Float64 restatement of the reference's cache precompute (codegen.cpp:254-292): Riccati fixed point
    on Q+rho, R+rho from P = rho*I, stop when max|dK| < 1e-5 (no failure if it never converges).
    Python-side helper for building synthetic problems; the product's tiny_precompute is C++."""
    A = np.asarray(A, np.float64); B = np.asarray(B, np.float64)
    n, m = B.shape
    Q1 = np.diag(np.asarray(Q, np.float64).reshape(-1) + rho)
    R1 = np.diag(np.asarray(R, np.float64).reshape(-1) + rho)
    Kt = np.zeros((m, n)); Pt = rho * np.eye(n)
    K = Kt; P = Pt; iters = max_riccati_iter
    for i in range(max_riccati_iter):
        K = np.linalg.inv(R1 + B.T @ Pt @ B) @ B.T @ Pt @ A
        P = Q1 + A.T @ Pt @ (A - B @ K)
        if np.abs(K - Kt).max() < tol:
            iters = i + 1
            break
        Kt, Pt = K, P
    Quu_inv = np.linalg.inv(R1 + B.T @ P @ B)
    AmBKt = (A - B @ K).T
    return {"Kinf": K, "Pinf": P, "Quu_inv": Quu_inv, "AmBKt": AmBKt, "riccati_iters": iters,
            "Q_rho": np.diag(Q1).copy()}


def cartpole(N: int = 10, max_iter: int = 100) -> Problem:
    """Config 4: cartpole model of examples/codegen_cartpole.cpp:22-28, cache by the codegen recursion,
    work.Q = Q + rho as tiny_codegen emits it (codegen.cpp:255,433), bounds +-5, tol 1e-3."""
    s, m = cartpole_model()
    rho = float(s["rho"])
    c = precompute_cache(m["Adyn"], m["Bdyn"], m["Q"], m["R"], rho)
    p = Problem(nx=4, nu=1, N=N, rho=rho, Adyn=m["Adyn"], Bdyn=m["Bdyn"], Q=c["Q_rho"], R=m["R"].reshape(-1),
                Kinf=c["Kinf"], Pinf=c["Pinf"], Quu_inv=c["Quu_inv"], AmBKt=c["AmBKt"], max_iter=max_iter,
                name="cartpole", extra={"riccati_iters": c["riccati_iters"]})
    return p.with_bounds(-5.0, 5.0, -5.0, 5.0)


def random_system(nx: int = 32, nu: int = 8, N: int = 50, seed: int = 2024, rho: float = 1.0) -> Problem:
    """Config 5: a seeded, controllable, mildly unstable random LTI system (the reference's codegen_random
    example is a 2x2 toy with infeasible bounds).  A = I + 0.05*G (G ~ U(-1,1), scaled to
    spectral radius <= 1.02), B ~ 0.2*U(-1,1); Q = 10, R = 1; |u| <= 1, |x| <= 10."""
    from .workloads import u01
    g = 2.0 * u01(seed, np.arange(nx * nx, dtype=np.uint64)).reshape(nx, nx) - 1.0
    A = np.eye(nx) + 0.05 * g
    A *= min(1.0, 1.02 / np.abs(np.linalg.eigvals(A)).max())
    Bm = 0.2 * (2.0 * u01(seed + 1, np.arange(nx * nu, dtype=np.uint64)).reshape(nx, nu) - 1.0)
    Q = np.full(nx, 10.0); R = np.full(nu, 1.0)
    c = precompute_cache(A, Bm, Q, R, rho)
    p = Problem(nx=nx, nu=nu, N=N, rho=rho, Adyn=A, Bdyn=Bm, Q=Q, R=R, Kinf=c["Kinf"], Pinf=c["Pinf"],
                Quu_inv=c["Quu_inv"], AmBKt=c["AmBKt"], name="random_%d_%d_%d" % (nx, nu, N),
                extra={"riccati_iters": c["riccati_iters"]})
    return p.with_bounds(-10.0, 10.0, -1.0, 1.0)


def codegen_random() -> Problem:
    """The reference's examples/codegen_random.cpp:20-39 verbatim: nx = 2, nu = 2, N = 3, column-major A = {1,5,1,2},
    B = {3,3,4,1}, Q = 1, R = 2, rho = 0.1 and its (inverted: min > max) box bounds; cache by the codegen recursion,
    work.Q = Q + rho as tiny_codegen stores it (codegen.cpp:255,433)."""
    A = np.array([1, 5, 1, 2], np.float64).reshape(2, 2, order="F")
    Bm = np.array([3, 3, 4, 1], np.float64).reshape(2, 2, order="F")
    Q = np.array([1.0, 1.0]); R = np.array([2.0, 2.0]); rho = 0.1
    c = precompute_cache(A, Bm, Q, R, rho)
    p = Problem(nx=2, nu=2, N=3, rho=rho, Adyn=A, Bdyn=Bm, Q=c["Q_rho"], R=R, Kinf=c["Kinf"], Pinf=c["Pinf"],
                Quu_inv=c["Quu_inv"], AmBKt=c["AmBKt"], name="codegen_random", extra={"riccati_iters": c["riccati_iters"]})
    p.x_min = np.tile(np.array([1.0, 2.0]), (3, 1)); p.x_max = np.tile(np.array([-1.0, -2.0]), (3, 1))
    p.u_min = np.tile(np.array([2.0, 3.0]), (2, 1)); p.u_max = np.tile(np.array([-2.0, -3.0]), (2, 1))
    return p
