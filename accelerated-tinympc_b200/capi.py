"""ctypes binding of the C ABI in include/tmpc.h (lib/libtmpc_cuda.so).

This is what a foreign-language caller of the reference would bind (the reference's own FFI surface is
tiny_wrapper.hpp:14-23, one global instance; here it is batched).  No torch types: raw addresses only.
The library has no CPU fallback: `load()` raises if the shared object is missing, `Solver()` raises if
no B200-class device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB_PATH = os.environ.get("TMPC_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtmpc_cuda.so")

TMPC_F32, TMPC_F64 = 0, 1
TMPC_ORDER_PARITY, TMPC_ORDER_FAST = 0, 1
TMPC_MEM_HOST, TMPC_MEM_DEVICE = 0, 1

EXPORTS = ["tmpc_create", "tmpc_destroy", "tmpc_set_model", "tmpc_set_settings", "tmpc_set_instance_bounds", "tmpc_solve", "tmpc_get_stats", "tmpc_step",
           "tmpc_host_alloc", "tmpc_host_free", "tmpc_last_error", "tmpc_version",
           "tmpc_batch_create", "tmpc_batch_destroy", "tmpc_batch_set_x0", "tmpc_batch_set_xref", "tmpc_batch_set_xref_table",
           "tmpc_batch_reset_dual_variables", "tmpc_batch_reset", "tmpc_batch_solve", "tmpc_batch_get", "tmpc_batch_rollout",
           "tmpc_batch_last_rollout_ms", "tmpc_batch_last_error",
           "tmpc_systems_precompute", "tmpc_systems_destroy", "tmpc_systems_get", "tmpc_solve_systems",
           "tmpc_multi_create", "tmpc_multi_destroy", "tmpc_multi_device_count", "tmpc_multi_ctx", "tmpc_multi_set_model",
           "tmpc_multi_set_settings", "tmpc_multi_set_instance_bounds", "tmpc_multi_solve", "tmpc_multi_rollout", "tmpc_multi_get_stats",
           "tmpc_multi_last_error", "tmpc_device_count"]

SYS = {"Kinf": 0, "Pinf": 1, "Quu_inv": 2, "AmBKt": 3, "Adyn": 4, "Bdyn": 5, "Q": 6, "rho": 7, "sweeps": 8}

GET = {"x": 0, "u": 1, "iter": 2, "status": 3, "resid": 4, "x0": 5, "d": 6, "y": 7, "z": 8, "g": 9, "v": 10}


class TmpcWarm(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d", "y", "g", "v", "z")]


class TmpcSolveArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("x0", C.c_void_p), ("Xref", C.c_void_p), ("xref_shared", C.c_int32),
                ("mem", C.c_int32), ("warm", C.POINTER(TmpcWarm)), ("x", C.c_void_p), ("u", C.c_void_p),
                ("iter", C.c_void_p), ("status", C.c_void_p), ("resid", C.c_void_p), ("stream", C.c_void_p),
                ("u0", C.c_void_p)]


class TmpcWorkspace(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x", "u", "q", "r", "p", "d", "v", "vnew", "z", "znew", "g", "y", "Xref")] + \
               [("xref_shared", C.c_int32), ("resid", C.c_void_p), ("term", C.c_void_p)]


class TmpcStats(C.Structure):
    _fields_ = [("instances", C.c_int64), ("iterations", C.c_int64), ("solved", C.c_int64), ("trips", C.c_int64),
                ("launches", C.c_int32), ("lanes", C.c_int32), ("kernel_ms", C.c_float), ("parity_pinned", C.c_int32),
                ("pattern", C.c_int32), ("scheduled", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with `make -C accelerated-tinympc_b200` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.tmpc_create.restype = C.c_int
    lib.tmpc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.tmpc_destroy.restype = C.c_int
    lib.tmpc_destroy.argtypes = [C.c_void_p]
    lib.tmpc_set_model.restype = C.c_int
    lib.tmpc_set_model.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.c_double] + [C.c_void_p] * 4
    lib.tmpc_set_settings.restype = C.c_int
    lib.tmpc_set_settings.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.tmpc_set_instance_bounds.restype = C.c_int
    lib.tmpc_set_instance_bounds.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int32]
    lib.tmpc_solve.restype = C.c_int
    lib.tmpc_solve.argtypes = [C.c_void_p, C.POINTER(TmpcSolveArgs)]
    lib.tmpc_get_stats.restype = C.c_int
    lib.tmpc_get_stats.argtypes = [C.c_void_p, C.POINTER(TmpcStats)]
    lib.tmpc_step.restype = C.c_int
    lib.tmpc_step.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(TmpcWorkspace), C.c_int32, C.c_int32, C.c_void_p]
    lib.tmpc_host_alloc.restype = C.c_int
    lib.tmpc_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    lib.tmpc_host_free.restype = C.c_int
    lib.tmpc_host_free.argtypes = [C.c_void_p]
    lib.tmpc_last_error.restype = C.c_char_p
    lib.tmpc_last_error.argtypes = [C.c_void_p]
    lib.tmpc_version.restype = C.c_char_p
    lib.tmpc_batch_create.restype = C.c_int
    lib.tmpc_batch_create.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    lib.tmpc_batch_destroy.restype = C.c_int
    lib.tmpc_batch_destroy.argtypes = [C.c_void_p]
    lib.tmpc_batch_set_x0.restype = C.c_int
    lib.tmpc_batch_set_x0.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.tmpc_batch_set_xref.restype = C.c_int
    lib.tmpc_batch_set_xref.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.tmpc_batch_set_xref_table.restype = C.c_int
    lib.tmpc_batch_set_xref_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]
    for fn in ("tmpc_batch_reset_dual_variables", "tmpc_batch_reset", "tmpc_batch_solve"):
        getattr(lib, fn).restype = C.c_int
        getattr(lib, fn).argtypes = [C.c_void_p]
    lib.tmpc_batch_get.restype = C.c_int
    lib.tmpc_batch_get.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    lib.tmpc_batch_rollout.restype = C.c_int
    lib.tmpc_batch_rollout.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.tmpc_batch_last_rollout_ms.restype = C.c_float
    lib.tmpc_batch_last_rollout_ms.argtypes = [C.c_void_p]
    lib.tmpc_batch_last_error.restype = C.c_char_p
    lib.tmpc_batch_last_error.argtypes = [C.c_void_p]
    lib.tmpc_systems_precompute.restype = C.c_int
    lib.tmpc_systems_precompute.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.tmpc_systems_destroy.restype = C.c_int
    lib.tmpc_systems_destroy.argtypes = [C.c_void_p]
    lib.tmpc_systems_get.restype = C.c_int
    lib.tmpc_systems_get.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.tmpc_solve_systems.restype = C.c_int
    lib.tmpc_solve_systems.argtypes = [C.c_void_p, C.POINTER(TmpcSolveArgs), C.c_void_p]
    lib.tmpc_multi_create.restype = C.c_int
    lib.tmpc_multi_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.tmpc_multi_destroy.restype = C.c_int
    lib.tmpc_multi_destroy.argtypes = [C.c_void_p]
    lib.tmpc_multi_device_count.restype = C.c_int
    lib.tmpc_multi_device_count.argtypes = [C.c_void_p]
    lib.tmpc_multi_ctx.restype = C.c_void_p
    lib.tmpc_multi_ctx.argtypes = [C.c_void_p, C.c_int]
    lib.tmpc_multi_set_model.restype = C.c_int
    lib.tmpc_multi_set_model.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.c_double] + [C.c_void_p] * 4
    lib.tmpc_multi_set_settings.restype = C.c_int
    lib.tmpc_multi_set_settings.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.tmpc_multi_set_instance_bounds.restype = C.c_int
    lib.tmpc_multi_set_instance_bounds.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4
    lib.tmpc_multi_solve.restype = C.c_int
    lib.tmpc_multi_solve.argtypes = [C.c_void_p, C.POINTER(TmpcSolveArgs)]
    lib.tmpc_multi_rollout.restype = C.c_int
    lib.tmpc_multi_rollout.argtypes = [C.c_void_p, C.c_void_p]
    lib.tmpc_multi_get_stats.restype = C.c_int
    lib.tmpc_multi_get_stats.argtypes = [C.c_void_p, C.POINTER(TmpcStats), C.c_void_p]
    lib.tmpc_multi_last_error.restype = C.c_char_p
    lib.tmpc_multi_last_error.argtypes = [C.c_void_p]
    _lib = lib
    return lib


class TmpcError(RuntimeError):
    pass


def make_solve_args(batch, x0, Xref, xref_shared, mem, x=None, u=None, it=None, status=None, resid=None, warm=None,
                    stream=None, u0=None):
    """tmpc_solve_args from address providers (numpy / torch / int / None).  Returns (args, keep-alive)."""
    args = TmpcSolveArgs()
    args.batch = batch
    args.x0 = _addr(x0)
    args.Xref = _addr(Xref)
    args.xref_shared = 1 if xref_shared else 0
    args.mem = mem
    w = None
    if warm is not None:
        w = TmpcWarm()
        for k in ("d", "y", "g", "v", "z"):
            setattr(w, k, _addr(warm[k]))
        args.warm = C.pointer(w)
    args.x, args.u, args.iter, args.status, args.resid = _addr(x), _addr(u), _addr(it), _addr(status), _addr(resid)
    args.u0 = _addr(u0)
    args.stream = stream
    return args, w


def host_io(s, x0, Xref, warm, outputs):
    """numpy inputs checked / cast for a solver-like object `s` (dtype, nx, nu, N) + empty output arrays for `outputs`."""
    dt = s.dtype
    x0 = np.ascontiguousarray(x0, dtype=dt).reshape(-1, s.nx)
    B = x0.shape[0]
    Xref = np.ascontiguousarray(Xref, dtype=dt)
    shared = Xref.size == s.N * s.nx
    if not shared and Xref.size != B * s.N * s.nx:
        raise ValueError("Xref must be [N,nx] or [B,N,nx]")
    shapes = {"x": ((B, s.N, s.nx), dt), "u": ((B, s.N - 1, s.nu), dt), "u0": ((B, s.nu), dt),
              "iter": ((B,), np.int32), "status": ((B,), np.int32), "resid": ((B, 4), dt)}
    out = {k: np.empty(*shapes[k]) for k in outputs}
    if warm is not None:
        warm = {k: np.ascontiguousarray(warm[k], dtype=dt) for k in ("d", "y", "g", "v", "z")}
        out["warm"] = warm
    return x0, Xref, shared, out, warm


def _addr(a):
    """Address of a numpy array, a torch tensor, an int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


class Solver:
    """One tmpc_ctx: a problem shape + model on one device.  Mirrors TinySolver{settings, cache, work}
    (types.hpp:102-107) for a batch."""

    def __init__(self, prob, dtype=np.float32, policy="parity", device=0):
        self.lib = load()
        self.prob = prob
        self.dtype = np.dtype(dtype)
        self.nx, self.nu, self.N = prob.nx, prob.nu, prob.N
        self._ctx = C.c_void_p()
        pol = {"parity": TMPC_ORDER_PARITY, "fast": TMPC_ORDER_FAST}[policy]
        rc = self.lib.tmpc_create(C.byref(self._ctx), device, prob.nx, prob.nu, prob.N,
                                  TMPC_F32 if self.dtype == np.float32 else TMPC_F64, pol)
        if rc != 0:
            raise TmpcError("tmpc_create: %d %s" % (rc, self.lib.tmpc_last_error(None).decode()))
        self.set_model(prob)

    def _check(self, rc, what):
        if rc != 0:
            raise TmpcError("%s: %d %s" % (what, rc, self.lib.tmpc_last_error(self._ctx).decode()))

    def set_model(self, prob):
        a = prob.cast(self.dtype)
        self._model_keep = a
        p = lambda k: _addr(a[k])
        self._check(self.lib.tmpc_set_model(self._ctx, p("Kinf"), p("Pinf"), p("Quu_inv"), p("AmBKt"), p("Adyn"),
                                            p("Bdyn"), p("Q"), float(prob.rho), p("x_min"), p("x_max"), p("u_min"),
                                            p("u_max")), "tmpc_set_model")
        self._check(self.lib.tmpc_set_settings(self._ctx, prob.abs_pri_tol, prob.abs_dua_tol, prob.max_iter,
                                               prob.check_termination, prob.en_state_bound, prob.en_input_bound),
                    "tmpc_set_settings")

    def set_instance_bounds(self, x_min=None, x_max=None, u_min=None, u_max=None):
        """Per-instance boxes (the wrapper's set_xmin ... set_umax, tiny_wrapper.cpp:43-129, with a batch dimension):
        numpy arrays [B, N, nx] / [B, N-1, nu]; all None = back to the shared bounds of the model."""
        if x_min is None and x_max is None and u_min is None and u_max is None:
            self._check(self.lib.tmpc_set_instance_bounds(self._ctx, 0, None, None, None, None, TMPC_MEM_HOST),
                        "tmpc_set_instance_bounds")
            return
        dt = self.dtype
        B = np.asarray(x_min).shape[0]
        a = [np.ascontiguousarray(v, dtype=dt) for v in (x_min, x_max, u_min, u_max)]
        for v, n in zip(a, (self.N * self.nx, self.N * self.nx, (self.N - 1) * self.nu, (self.N - 1) * self.nu)):
            if v.size != B * n:
                raise ValueError("bounds must be [B,N,nx] / [B,N-1,nu]")
        self._check(self.lib.tmpc_set_instance_bounds(self._ctx, B, _addr(a[0]), _addr(a[1]), _addr(a[2]), _addr(a[3]),
                                                      TMPC_MEM_HOST), "tmpc_set_instance_bounds")

    def solve_raw(self, batch, x0, Xref, xref_shared, mem, x=None, u=None, it=None, status=None, resid=None,
                  warm=None, stream=None, u0=None):
        """Thin call: every array argument is an address provider (numpy / torch / int) or None."""
        args, keep = make_solve_args(batch, x0, Xref, xref_shared, mem, x, u, it, status, resid, warm, stream, u0)
        self._check(self.lib.tmpc_solve(self._ctx, C.byref(args)), "tmpc_solve")

    def solve(self, x0, Xref, warm=None, outputs=("x", "u", "iter", "status", "resid")):
        """Host-array convenience (numpy in, numpy out) through TMPC_MEM_HOST.  `outputs` is the output mask: any of
        x, u, u0 (= u(:,0) alone), iter, status, resid."""
        x0, Xref, shared, out, warm = host_io(self, x0, Xref, warm, outputs)
        self.solve_raw(x0.shape[0], x0, Xref, shared, TMPC_MEM_HOST, out.get("x"), out.get("u"), out.get("iter"),
                       out.get("status"), out.get("resid"), warm=warm, u0=out.get("u0"))
        return out

    def step(self, which, ws, it=1):
        """One reference step function on host workspaces.  `ws`: dict of numpy arrays x,u,q,r,p,d,v,vnew,z,znew,g,y
        ([B,N,nx] / [B,N-1,nu]), Xref ([N,nx] or [B,N,nx]), resid [B,4]; modified in place.  Returns term[B]."""
        dt = self.dtype
        B = ws["x"].shape[0]
        w = TmpcWorkspace()
        for k in ("x", "u", "q", "r", "p", "d", "v", "vnew", "z", "znew", "g", "y", "Xref", "resid"):
            a = ws[k]
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], k
            setattr(w, k, _addr(a))
        w.xref_shared = 1 if ws["Xref"].size == self.N * self.nx else 0
        term = np.zeros(B, np.int32)
        w.term = _addr(term)
        self._check(self.lib.tmpc_step(self._ctx, which, B, C.byref(w), it, TMPC_MEM_HOST, None), "tmpc_step")
        return term

    def stats(self):
        s = TmpcStats()
        self._check(self.lib.tmpc_get_stats(self._ctx, C.byref(s)), "tmpc_get_stats")
        return {f[0]: getattr(s, f[0]) for f in TmpcStats._fields_}

    def close(self):
        if self._ctx:
            self.lib.tmpc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RolloutArgs(C.Structure):
    """tmpc_rollout_args (include/tmpc.h)."""
    _fields_ = [("batch", C.c_int64), ("steps", C.c_int32), ("reset_duals", C.c_int32), ("x0", C.c_void_p), ("Xref", C.c_void_p),
                ("xref_shared", C.c_int32), ("table", C.c_void_p), ("table_rows", C.c_int64), ("start", C.POINTER(C.c_int32)),
                ("x0_hist", C.c_void_p), ("u0_hist", C.c_void_p), ("iter_hist", C.POINTER(C.c_int32)), ("status_hist", C.POINTER(C.c_int32)),
                ("x", C.c_void_p), ("u", C.c_void_p)]


class Multi:
    """tmpc_multi: one host batch over several devices from one process (one ctx + one host worker thread per device,
    contiguous instance ranges, no inter-device traffic).  devices: None = every visible device, an int n = the first n,
    or a list of device indices."""

    def __init__(self, prob, dtype=np.float32, policy="parity", devices=None):
        self.lib = load()
        self.prob = prob
        self.dtype = np.dtype(dtype)
        self.nx, self.nu, self.N = prob.nx, prob.nu, prob.N
        self._m = C.c_void_p()
        pol = {"parity": TMPC_ORDER_PARITY, "fast": TMPC_ORDER_FAST}[policy]
        if devices is None:
            n, arr = 0, None
        elif isinstance(devices, int):
            n, arr = devices, None
        else:
            n, arr = len(devices), (C.c_int * len(devices))(*devices)
        rc = self.lib.tmpc_multi_create(C.byref(self._m), n, arr, prob.nx, prob.nu, prob.N,
                                        TMPC_F32 if self.dtype == np.float32 else TMPC_F64, pol)
        if rc != 0:
            raise TmpcError("tmpc_multi_create: %d %s" % (rc, self.lib.tmpc_multi_last_error(None).decode()))
        a = prob.cast(self.dtype)
        self._model_keep = a
        p = lambda k: _addr(a[k])
        self._check(self.lib.tmpc_multi_set_model(self._m, p("Kinf"), p("Pinf"), p("Quu_inv"), p("AmBKt"), p("Adyn"), p("Bdyn"),
                                                  p("Q"), float(prob.rho), p("x_min"), p("x_max"), p("u_min"), p("u_max")),
                    "tmpc_multi_set_model")
        self._check(self.lib.tmpc_multi_set_settings(self._m, prob.abs_pri_tol, prob.abs_dua_tol, prob.max_iter,
                                                     prob.check_termination, prob.en_state_bound, prob.en_input_bound),
                    "tmpc_multi_set_settings")

    def _check(self, rc, what):
        if rc != 0:
            raise TmpcError("%s: %d %s" % (what, rc, self.lib.tmpc_multi_last_error(self._m).decode()))

    @property
    def device_count(self):
        return int(self.lib.tmpc_multi_device_count(self._m))

    def set_instance_bounds(self, x_min=None, x_max=None, u_min=None, u_max=None):
        if x_min is None:
            return self._check(self.lib.tmpc_multi_set_instance_bounds(self._m, 0, None, None, None, None), "tmpc_multi_set_instance_bounds")
        a = [np.ascontiguousarray(v, dtype=self.dtype) for v in (x_min, x_max, u_min, u_max)]
        self._ib_keep = a
        self._check(self.lib.tmpc_multi_set_instance_bounds(self._m, a[0].shape[0], _addr(a[0]), _addr(a[1]), _addr(a[2]), _addr(a[3])),
                    "tmpc_multi_set_instance_bounds")

    def solve_raw(self, batch, x0, Xref, xref_shared, x=None, u=None, it=None, status=None, resid=None, warm=None, u0=None):
        args, keep = make_solve_args(batch, x0, Xref, xref_shared, TMPC_MEM_HOST, x, u, it, status, resid, warm, None, u0)
        self._check(self.lib.tmpc_multi_solve(self._m, C.byref(args)), "tmpc_multi_solve")

    def solve(self, x0, Xref, warm=None, outputs=("x", "u", "iter", "status", "resid")):
        x0, Xref, shared, out, warm = host_io(self, x0, Xref, warm, outputs)
        self.solve_raw(x0.shape[0], x0, Xref, shared, out.get("x"), out.get("u"), out.get("iter"), out.get("status"),
                       out.get("resid"), warm=warm, u0=out.get("u0"))
        return out

    def rollout(self, x0, steps, xref=None, table=None, start=None, reset_duals=True, last=False):
        """tmpc_multi_rollout: the closed loop of every instance on the device that owns its index range.  Returns the histories
        (x0 [steps+1,B,nx], u0 [steps,B,nu], iter, status [steps,B]) and, with last=True, x / u of the last solve."""
        x0 = np.ascontiguousarray(x0, dtype=self.dtype)
        B = x0.shape[0]
        r = RolloutArgs()
        keep = [x0]
        r.batch, r.steps, r.reset_duals, r.x0 = B, steps, 1 if reset_duals else 0, _addr(x0)
        if table is not None:
            t = np.ascontiguousarray(table, dtype=self.dtype).reshape(-1, self.nx)
            keep.append(t)
            r.table, r.table_rows = _addr(t), t.shape[0]
            if start is not None:
                st = np.ascontiguousarray(start, dtype=np.int32).reshape(B)
                keep.append(st)
                r.start = st.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            xr = np.ascontiguousarray(xref, dtype=self.dtype)
            keep.append(xr)
            r.Xref, r.xref_shared = _addr(xr), 1 if xr.ndim == 2 else 0
        h = {"x0": np.empty((steps + 1, B, self.nx), self.dtype), "u0": np.empty((steps, B, self.nu), self.dtype),
             "iter": np.empty((steps, B), np.int32), "status": np.empty((steps, B), np.int32)}
        r.x0_hist, r.u0_hist = _addr(h["x0"]), _addr(h["u0"])
        r.iter_hist = h["iter"].ctypes.data_as(C.POINTER(C.c_int32))
        r.status_hist = h["status"].ctypes.data_as(C.POINTER(C.c_int32))
        if last:
            h["x"] = np.empty((B, self.N, self.nx), self.dtype)
            h["u"] = np.empty((B, self.N - 1, self.nu), self.dtype)
            r.x, r.u = _addr(h["x"]), _addr(h["u"])
        self._check(self.lib.tmpc_multi_rollout(self._m, C.byref(r)), "tmpc_multi_rollout")
        return h

    def stats(self):
        n = self.device_count
        tot, per = TmpcStats(), (TmpcStats * n)()
        self._check(self.lib.tmpc_multi_get_stats(self._m, C.byref(tot), per), "tmpc_multi_get_stats")
        d = {f[0]: getattr(tot, f[0]) for f in TmpcStats._fields_}
        d["per_device"] = [{f[0]: getattr(per[i], f[0]) for f in TmpcStats._fields_} for i in range(n)]
        return d

    def close(self):
        if self._m:
            self.lib.tmpc_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """tmpc_batch: device-resident workspaces of `batch` instances of a Solver -- the reference's wrapper calls
    (tiny_wrapper.hpp:14-23) with a leading batch dimension, and the examples' closed loop on the device."""

    def __init__(self, solver: Solver, batch: int):
        self.s = solver
        self.lib = solver.lib
        self.B = int(batch)
        self._b = C.c_void_p()
        solver._check(self.lib.tmpc_batch_create(solver._ctx, self.B, C.byref(self._b)), "tmpc_batch_create")

    def _check(self, rc, what):
        if rc != 0:
            raise TmpcError("%s: %d %s" % (what, rc, self.lib.tmpc_batch_last_error(self._b).decode()))

    def _host(self, a, shape=None):
        a = np.ascontiguousarray(a, dtype=self.s.dtype)
        return a if shape is None else a.reshape(shape)

    def set_x0(self, x0):
        if hasattr(x0, "data_ptr"):
            return self._check(self.lib.tmpc_batch_set_x0(self._b, x0.data_ptr(), TMPC_MEM_DEVICE), "tmpc_batch_set_x0")
        a = self._host(x0, (self.B, self.s.nx))
        self._check(self.lib.tmpc_batch_set_x0(self._b, a.ctypes.data, TMPC_MEM_HOST), "tmpc_batch_set_x0")

    def set_xref(self, xref):
        a = self._host(xref)
        shared = a.size == self.s.N * self.s.nx
        if not shared and a.size != self.B * self.s.N * self.s.nx:
            raise ValueError("Xref must be [N,nx] or [B,N,nx]")
        self._check(self.lib.tmpc_batch_set_xref(self._b, a.ctypes.data, 1 if shared else 0, TMPC_MEM_HOST), "tmpc_batch_set_xref")

    def set_xref_table(self, table, start=None):
        t = self._host(table).reshape(-1, self.s.nx)
        st = None if start is None else np.ascontiguousarray(start, dtype=np.int32).reshape(self.B)
        self._check(self.lib.tmpc_batch_set_xref_table(self._b, t.ctypes.data, t.shape[0], None if st is None else st.ctypes.data,
                                                       TMPC_MEM_HOST), "tmpc_batch_set_xref_table")

    def reset_dual_variables(self):
        self._check(self.lib.tmpc_batch_reset_dual_variables(self._b), "tmpc_batch_reset_dual_variables")

    def reset(self):
        self._check(self.lib.tmpc_batch_reset(self._b), "tmpc_batch_reset")

    def solve(self):
        self._check(self.lib.tmpc_batch_solve(self._b), "tmpc_batch_solve")

    def get(self, what):
        s = self.s
        shape = {"x": (self.B, s.N, s.nx), "u": (self.B, s.N - 1, s.nu), "iter": (self.B,), "status": (self.B,), "resid": (self.B, 4),
                 "x0": (self.B, s.nx), "d": (self.B, s.N - 1, s.nu), "y": (self.B, s.N - 1, s.nu), "z": (self.B, s.N - 1, s.nu),
                 "g": (self.B, s.N, s.nx), "v": (self.B, s.N, s.nx)}[what]
        out = np.empty(shape, np.int32 if what in ("iter", "status") else s.dtype)
        self._check(self.lib.tmpc_batch_get(self._b, GET[what], out.ctypes.data, TMPC_MEM_HOST), "tmpc_batch_get")
        return out

    def rollout(self, steps, reset_duals=True, history=True):
        """`steps` closed-loop MPC steps on the device.  Returns dict(x0 [steps+1,B,nx], u0 [steps,B,nu], iter, status
        [steps,B]) when history is requested, else None."""
        s = self.s
        if not history:
            self._check(self.lib.tmpc_batch_rollout(self._b, steps, 1 if reset_duals else 0, None, None, None, None, TMPC_MEM_HOST),
                        "tmpc_batch_rollout")
            return None
        h = {"x0": np.empty((steps + 1, self.B, s.nx), s.dtype), "u0": np.empty((steps, self.B, s.nu), s.dtype),
             "iter": np.empty((steps, self.B), np.int32), "status": np.empty((steps, self.B), np.int32)}
        self._check(self.lib.tmpc_batch_rollout(self._b, steps, 1 if reset_duals else 0, h["x0"].ctypes.data, h["u0"].ctypes.data,
                                                h["iter"].ctypes.data, h["status"].ctypes.data, TMPC_MEM_HOST), "tmpc_batch_rollout")
        return h

    def last_rollout_ms(self):
        return float(self.lib.tmpc_batch_last_rollout_ms(self._b))

    def close(self):
        if self._b:
            self.lib.tmpc_batch_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Systems:
    """tmpc_systems: per-instance models + caches, precomputed on the device (codegen.cpp:254-292 per instance)."""

    def __init__(self, solver: Solver, Adyn, Bdyn, Q, R, rho, q_plus_rho=False):
        """Adyn [B, nx, nx], Bdyn [B, nx, nu] (numpy, natural row/col indexing), Q [B, nx], R [B, nu], rho [B]."""
        self.s = solver
        self.lib = solver.lib
        dt = solver.dtype
        A = np.ascontiguousarray(np.transpose(np.asarray(Adyn, dtype=np.float64), (0, 2, 1))).astype(dt)   # column-major per instance
        Bm = np.ascontiguousarray(np.transpose(np.asarray(Bdyn, dtype=np.float64), (0, 2, 1))).astype(dt)
        self.B = A.shape[0]
        Qa = np.ascontiguousarray(Q, dtype=dt).reshape(self.B, solver.nx)
        Ra = np.ascontiguousarray(R, dtype=dt).reshape(self.B, solver.nu)
        rh = np.ascontiguousarray(rho, dtype=dt).reshape(self.B)
        self.inputs = {"Adyn": A, "Bdyn": Bm, "Q": Qa, "R": Ra, "rho": rh}
        self._p = C.c_void_p()
        solver._check(self.lib.tmpc_systems_precompute(solver._ctx, self.B, A.ctypes.data, Bm.ctypes.data, Qa.ctypes.data, Ra.ctypes.data,
                                                       rh.ctypes.data, 1 if q_plus_rho else 0, TMPC_MEM_HOST, C.byref(self._p)),
                      "tmpc_systems_precompute")

    def get(self, what):
        """Cache matrices as [B, rows, cols] numpy arrays in natural indexing; sweeps / rho / Q as [B(, nx)]."""
        s = self.s
        nx, nu = s.nx, s.nu
        shape = {"Kinf": (nu, nx), "Pinf": (nx, nx), "Quu_inv": (nu, nu), "AmBKt": (nx, nx), "Adyn": (nx, nx), "Bdyn": (nx, nu)}
        if what == "sweeps":
            out = np.empty(self.B, np.int32)
        elif what == "rho":
            out = np.empty(self.B, s.dtype)
        elif what == "Q":
            out = np.empty((self.B, nx), s.dtype)
        else:
            r, c = shape[what]
            out = np.empty((self.B, c, r), s.dtype)      # column-major per instance
        s._check(self.lib.tmpc_systems_get(self._p, SYS[what], out.ctypes.data), "tmpc_systems_get")
        return np.transpose(out, (0, 2, 1)) if what in shape else out

    def solve_raw(self, x0, Xref, xref_shared, x=None, u=None, it=None, status=None, resid=None, warm=None, stream=None, u0=None):
        """Device buffers (torch tensors / addresses) only."""
        args = TmpcSolveArgs()
        args.batch = self.B
        args.u0 = _addr(u0)
        args.x0, args.Xref = _addr(x0), _addr(Xref)
        args.xref_shared = 1 if xref_shared else 0
        args.mem = TMPC_MEM_DEVICE
        w = None
        if warm is not None:
            w = TmpcWarm()
            for k in ("d", "y", "g", "v", "z"):
                setattr(w, k, _addr(warm[k]))
            args.warm = C.pointer(w)
        args.x, args.u, args.iter, args.status, args.resid = _addr(x), _addr(u), _addr(it), _addr(status), _addr(resid)
        args.stream = stream
        self.s._check(self.lib.tmpc_solve_systems(self.s._ctx, C.byref(args), self._p), "tmpc_solve_systems")

    def close(self):
        if self._p:
            self.lib.tmpc_systems_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
