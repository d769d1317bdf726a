"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for the two CPU checkers.

  RefLib(cfg)   oracle/_ref/libref_<cfg>.so  -- the unmodified reference (ref_harness.cpp), one per shape/scalar
  OracleLib()   oracle/libtinympc_oracle.so  -- the plain-C restatement (tinympc_oracle.c), any shape

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

SHAPES = {"q": (12, 4, 10), "c": (4, 1, 10), "l": (32, 8, 50)}


class _RefProblem(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("Kinf", "Pinf", "Quu_inv", "AmBKt", "Adyn", "Bdyn", "Q", "x_min", "x_max", "u_min", "u_max")] + \
               [("rho", C.c_double), ("abs_pri_tol", C.c_double), ("abs_dua_tol", C.c_double),
                ("max_iter", C.c_int32), ("check_termination", C.c_int32),
                ("en_state_bound", C.c_int32), ("en_input_bound", C.c_int32)]


class _OraProblem(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("N", C.c_int32), ("scalar_bytes", C.c_int32)] + \
               list(_RefProblem._fields_)


class _State(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d", "y", "g", "v", "z", "vnew", "znew", "q", "r", "p")]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fill_problem(P, prob, dtype, keep):
    arrs = prob.cast(dtype)
    keep.append(arrs)
    for k in ("Kinf", "Pinf", "Quu_inv", "AmBKt", "Adyn", "Bdyn", "Q", "x_min", "x_max", "u_min", "u_max"):
        a = arrs[k]
        if a is None and k in ("x_min", "x_max", "u_min", "u_max"):
            # the reference reads the bound arrays only when the matching en_*_bound flag is set
            a = np.zeros(prob.nx * prob.N, dtype=dtype)
            keep.append(a)
        setattr(P, k, _ptr(a))
    P.rho = prob.rho
    P.abs_pri_tol = prob.abs_pri_tol
    P.abs_dua_tol = prob.abs_dua_tol
    P.max_iter = prob.max_iter
    P.check_termination = prob.check_termination
    P.en_state_bound = prob.en_state_bound
    P.en_input_bound = prob.en_input_bound


class Result:
    __slots__ = ("x", "u", "iter", "status", "resid", "state")


STATE_DIMS = {"d": "u", "y": "u", "z": "u", "znew": "u", "r": "u", "g": "x", "v": "x", "vnew": "x", "q": "x", "p": "x"}


class _Base:
    """Shared batch driver.  `warm` = dict with any of d,y,g,v,z ([B, N-1, nu] / [B, N, nx]); missing = zeros."""

    def _solve(self, call, P, prob, dtype, x0, Xref, warm, want_state, nthreads):
        nx, nu, N = prob.nx, prob.nu, prob.N
        x0 = np.ascontiguousarray(x0, dtype=dtype).reshape(-1, nx)
        B = x0.shape[0]
        Xref = np.ascontiguousarray(Xref, dtype=dtype)
        if Xref.size == N * nx:
            stride = 0
        elif Xref.size == B * N * nx:
            stride = N * nx
        else:
            raise ValueError("Xref must be [N,nx] or [B,N,nx]")
        r = Result()
        r.x = np.empty((B, N, nx), dtype)
        r.u = np.empty((B, N - 1, nu), dtype)
        r.iter = np.empty(B, np.int32)
        r.status = np.empty(B, np.int32)
        r.resid = np.empty((B, 4), dtype)
        S = None
        r.state = None
        if warm is not None or want_state:
            S = _State()
            r.state = {}
            for k, kind in STATE_DIMS.items():
                shape = (B, N - 1, nu) if kind == "u" else (B, N, nx)
                if warm is not None and k in warm and k in ("d", "y", "g", "v", "z"):
                    a = np.array(warm[k], dtype=dtype, order="C").reshape(shape)
                else:
                    a = np.zeros(shape, dtype)
                r.state[k] = a
                setattr(S, k, _ptr(a))
        rc = call(P, B, x0, Xref, stride, S, r, nthreads)
        if rc != 0:
            raise RuntimeError("oracle call failed: %d" % rc)
        return r


class RefLib(_Base):
    def __init__(self, cfg: str):
        path = os.path.join(REF_DIR, "libref_%s.so" % cfg)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.cfg = cfg
        self.lib = C.CDLL(path)
        nx, nu, N, sb = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self.lib.ref_info(C.byref(nx), C.byref(nu), C.byref(N), C.byref(sb))
        self.nx, self.nu, self.N = nx.value, nu.value, N.value
        self.dtype = np.float32 if sb.value == 4 else np.float64
        self.lib.ref_solve_batch.restype = C.c_int
        self.lib.ref_solve_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int32]
        self.lib.ref_step.restype = C.c_int
        self.lib.ref_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        self.lib.ref_plant_step.restype = C.c_int
        self.lib.ref_plant_step.argtypes = [C.c_void_p] * 6

    @staticmethod
    def available(cfg: str) -> bool:
        return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % cfg))

    def _problem(self, prob):
        assert (prob.nx, prob.nu, prob.N) == (self.nx, self.nu, self.N), "shape mismatch with %s" % self.cfg
        keep = []
        P = _RefProblem()
        _fill_problem(P, prob, self.dtype, keep)
        return P, keep

    def solve_batch(self, prob, x0, Xref, warm=None, want_state=False, nthreads=1):
        P, keep = self._problem(prob)

        def call(P, B, x0, Xref, stride, S, r, nthreads):
            return self.lib.ref_solve_batch(C.byref(P), B, _ptr(x0), _ptr(Xref), stride,
                                            C.byref(S) if S is not None else None, _ptr(r.x), _ptr(r.u),
                                            _ptr(r.iter), _ptr(r.status), _ptr(r.resid), None, nthreads)
        return self._solve(call, P, prob, self.dtype, x0, Xref, warm, want_state, nthreads)

    def step(self, prob, which: int, ws: np.ndarray, it: int = 1):
        P, keep = self._problem(prob)
        ws = np.ascontiguousarray(ws, dtype=self.dtype).copy()
        rc = self.lib.ref_step(C.byref(P), which, _ptr(ws), it)
        return rc, ws

    def plant_step(self, prob, x0, u, xref=None):
        """The examples' `x1 = Adyn * x0 + Bdyn * u.col(0)` on the reference's Eigen types, one instance.
        u: [N-1, nu] trajectory.  Returns x1 (and the printed tracking error if xref [N, nx] is given)."""
        P, keep = self._problem(prob)
        x0 = np.ascontiguousarray(x0, dtype=self.dtype).reshape(-1)
        u = np.ascontiguousarray(u, dtype=self.dtype).reshape(self.N - 1, self.nu)
        x1 = np.zeros(self.nx, self.dtype)
        err = np.zeros(1, self.dtype)
        xr = None if xref is None else np.ascontiguousarray(xref, dtype=self.dtype)
        self.lib.ref_plant_step(C.byref(P), _ptr(x0), _ptr(u), _ptr(x1), _ptr(xr), _ptr(err) if xr is not None else None)
        return (x1, err[0]) if xr is not None else x1


class OracleLib(_Base):
    def __init__(self):
        path = os.path.join(HERE, "libtinympc_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run: make -C oracle oracle)")
        self.lib = C.CDLL(path)
        self.lib.oracle_solve_batch.restype = C.c_int
        self.lib.oracle_solve_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        self.lib.oracle_step.restype = C.c_int
        self.lib.oracle_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]

    def _problem(self, prob, dtype):
        keep = []
        P = _OraProblem()
        P.nx, P.nu, P.N = prob.nx, prob.nu, prob.N
        P.scalar_bytes = np.dtype(dtype).itemsize
        _fill_problem(P, prob, dtype, keep)
        return P, keep

    def solve_batch(self, prob, x0, Xref, dtype=np.float32, warm=None, want_state=False, nthreads=1):
        P, keep = self._problem(prob, dtype)

        def call(P, B, x0, Xref, stride, S, r, nthreads):
            return self.lib.oracle_solve_batch(C.byref(P), B, _ptr(x0), _ptr(Xref), stride,
                                               C.byref(S) if S is not None else None, _ptr(r.x), _ptr(r.u),
                                               _ptr(r.iter), _ptr(r.status), _ptr(r.resid), nthreads)
        return self._solve(call, P, prob, dtype, x0, Xref, warm, want_state, nthreads)

    def step(self, prob, which: int, ws: np.ndarray, it: int = 1, dtype=np.float32):
        P, keep = self._problem(prob, dtype)
        ws = np.ascontiguousarray(ws, dtype=dtype).copy()
        rc = self.lib.oracle_step(C.byref(P), which, _ptr(ws), it)
        return rc, ws

    def plant_step(self, prob, x0, u0, dtype=np.float32):
        """Batched plant step x1 = Adyn x0 + Bdyn u0 in the reference's evaluation order.  x0 [B, nx], u0 [B, nu]."""
        P, keep = self._problem(prob, dtype)
        x0 = np.ascontiguousarray(x0, dtype=dtype).reshape(-1, prob.nx)
        u0 = np.ascontiguousarray(u0, dtype=dtype).reshape(-1, prob.nu)
        x1 = np.empty_like(x0)
        self.lib.oracle_plant_step.restype = C.c_int
        self.lib.oracle_plant_step.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        rc = self.lib.oracle_plant_step(C.byref(P), x0.shape[0], _ptr(x0), _ptr(u0), prob.nu, _ptr(x1))
        if rc != 0:
            raise RuntimeError("oracle_plant_step failed")
        return x1


def ws_size(nx, nu, N):
    """Length of the step-function workspace image (see ref_step in ref_harness.cpp)."""
    return 7 * nx * N + 6 * nu * (N - 1) + 4
