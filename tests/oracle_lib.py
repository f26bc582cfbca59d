"""ctypes view of oracle/harness.h (TEST INFRASTRUCTURE: the checker, never the product path)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib_path(kind, problem, full_ddp, fast=False):
    if kind == "reference-mt":      # the reference's -DMULTI_THREADED=1 build (one solve at a time per process)
        return os.path.join(ROOT, "oracle", "_ref", f"libref_{problem}_ddp{int(full_ddp)}_mt.so")
    d = "_ref" if kind == "reference" else "_build"
    stem = {"reference": "libref", "port": "libport", "b200": "libb200h"}[kind]   # b200 = harness over the GPU drop-in library
    return os.path.join(ROOT, "oracle", d, f"{stem}_{problem}_ddp{int(full_ddp)}{'_fast' if fast else ''}.so")


def available(kind, problem, full_ddp, fast=False):
    return os.path.exists(lib_path(kind, problem, full_ddp, fast))


class OracleLib:
    def __init__(self, kind, problem, full_ddp=0, fast=False):
        self.path = lib_path(kind, problem, full_ddp, fast)
        L = self.lib = C.CDLL(self.path)
        L.h_create.restype = C.c_void_p
        L.h_create.argtypes = [C.c_int]
        L.h_destroy.argtypes = [C.c_void_p]
        L.h_set_opt.restype = C.c_char_p
        L.h_set_opt.argtypes = [C.c_void_p, C.c_char_p, _dp, C.c_int]
        L.h_set_param.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int]
        L.h_init.argtypes = [C.c_void_p, _dp, _dp]
        L.h_solve.argtypes = [C.c_void_p]
        for f in ("h_calc_derivs", "h_back_pass"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.h_line_search.argtypes = [C.c_void_p, C.c_int]
        L.h_forward_pass.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.c_int]
        L.h_make_candidate_nominal.argtypes = [C.c_void_p]
        L.h_update_multipliers.argtypes = [C.c_void_p, C.c_int]
        L.h_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.h_get.argtypes = [C.c_void_p, C.c_char_p, _dp]
        L.h_scalar.restype = C.c_double
        L.h_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.h_trace_len.argtypes = [C.c_void_p]
        L.h_trace.argtypes = [C.c_void_p, C.c_char_p, _dp]
        L.h_bp_trace_len.argtypes = [C.c_void_p]
        L.h_bp_trace.argtypes = [C.c_void_p, C.c_char_p, _dp]
        L.h_qp_trace_enable.argtypes = [C.c_void_p, C.c_int]
        L.h_qp_trace_len.argtypes = [C.c_void_p]
        L.h_qp_trace.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.h_param_name.restype = C.c_char_p
        L.h_kind.restype = C.c_char_p
        L.h_solve_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_char_p), _dp, C.c_int, C.c_int,
                                    _dp, _ip, _ip, _ip, C.c_void_p, C.c_void_p]
        self.nx, self.nu = L.h_nx(), L.h_nu()
        self.full_ddp = L.h_full_ddp()
        self.param_names = [L.h_param_name(i).decode() for i in range(L.h_n_params())]
        self.param_sizes = [L.h_param_size(i) for i in range(L.h_n_params())]
        self.kind = L.h_kind().decode()

    def solver(self, T):
        return OracleSolver(self, T)

    def flat_params(self, params, T):
        out = []
        for n, s in zip(self.param_names, self.param_sizes):
            v = np.asarray(params[n], dtype=np.float64).ravel()
            assert v.size == (T + 1 if s == -1 else s), n
            out.append(v)
        return np.ascontiguousarray(np.concatenate(out)) if out else np.zeros(1)

    def solve_batch(self, x0, u0, params, opts, n_threads, want_traj=False):
        B, T = u0.shape[0], u0.shape[1]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        names = (C.c_char_p * max(len(opts), 1))(*[k.encode() for k in opts])
        vals = np.ascontiguousarray(list(opts.values()) or [0.0], dtype=np.float64)
        cost = np.zeros(B)
        it = np.zeros(B, np.int32)
        nls = np.zeros(B, np.int32)
        res = np.zeros(B, np.int32)
        xo = np.zeros((B, T + 1, self.nx)) if want_traj else None
        uo = np.zeros((B, T, self.nu)) if want_traj else None
        self.lib.h_solve_batch(B, T, x0, u0, self.flat_params(params, T), names, vals, len(opts), n_threads,
                               cost, it, nls, res,
                               xo.ctypes.data_as(C.c_void_p) if want_traj else None,
                               uo.ctypes.data_as(C.c_void_p) if want_traj else None)
        return dict(cost=cost, iterations=it, n_linesearch=nls, result=res, x=xo, u=uo)


class OracleSolver:
    def __init__(self, olib, T):
        self.o = olib
        self.L = olib.lib
        self.T = T
        self.h = self.L.h_create(T)

    def close(self):
        if self.h:
            self.L.h_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_opts(self, opts):
        for k, v in opts.items():
            v = np.atleast_1d(np.asarray(v, dtype=np.float64))
            err = self.L.h_set_opt(self.h, k.encode(), np.ascontiguousarray(v), v.size)
            if err:
                raise ValueError(f"{k}: {err.decode()}")

    def set_opt_raw(self, name, v):
        v = np.atleast_1d(np.asarray(v, dtype=np.float64))
        err = self.L.h_set_opt(self.h, name.encode(), np.ascontiguousarray(v), v.size)
        return err.decode() if err else None

    def set_params(self, params):
        for i, n in enumerate(self.o.param_names):
            v = np.ascontiguousarray(np.asarray(params[n], dtype=np.float64).ravel())
            assert self.L.h_set_param(self.h, i, v, v.size), n

    def init(self, x0, u0):
        return self.L.h_init(self.h, np.ascontiguousarray(x0, dtype=np.float64),
                             np.ascontiguousarray(u0, dtype=np.float64))

    def solve(self):
        return self.L.h_solve(self.h)

    def calc_derivs(self):
        return self.L.h_calc_derivs(self.h)

    def back_pass(self):
        return self.L.h_back_pass(self.h)

    def line_search(self, it):
        return self.L.h_line_search(self.h, it)

    def forward_pass(self, alpha, cost_only=0):
        c = C.c_double(0.0)
        ok = self.L.h_forward_pass(self.h, alpha, C.byref(c), cost_only)
        return ok, c.value

    def make_candidate_nominal(self):
        self.L.h_make_candidate_nominal(self.h)

    def update_multipliers(self, init):
        return self.L.h_update_multipliers(self.h, init)

    def set_scalar(self, name, v):
        self.L.h_set_scalar(self.h, name.encode(), float(v))

    def scalar(self, name):
        return self.L.h_scalar(self.h, name.encode())

    _SHAPES = None

    def get(self, field):
        nx, nu, T = self.o.nx, self.o.nu, self.T
        nqxx, nquu = nx * (nx + 1) // 2, nu * (nu + 1) // 2
        shapes = {"x": (T + 1, nx), "u": (T, nu), "l": (T, nu), "L": (T, nu * nx), "lower": (T, nu), "upper": (T, nu),
                  "lower_sign": (T, nu), "upper_sign": (T, nu), "lower_hx": (T, nu * nx), "upper_hx": (T, nu * nx),
                  "fx": (T, nx * nx), "fu": (T, nx * nu), "cu": (T, nu), "cuu": (T, nquu), "cxu": (T, nx * nu),
                  "fxx": (T, nx * nqxx), "fuu": (T, nx * nquu), "fxu": (T, nx * nx * nu), "c": (T + 1,),
                  "cx": (T + 1, nx), "cxx": (T + 1, nqxx), "mult_f": (64,), "mult_t": (T, 64), "log_linesearch": (4096,)}
        buf = np.zeros(int(np.prod(shapes[field])) + 8)
        n = self.L.h_get(self.h, field.encode(), buf)
        if n < 0:
            raise KeyError(field)
        if field in ("mult_f", "log_linesearch"):
            return buf[:n].copy()
        if field == "mult_t":
            return buf[:n].reshape(T, -1).copy() if n else np.zeros((T, 0))
        return buf[:n].reshape(shapes[field]).copy()

    def trace(self, what):
        n = self.L.h_trace_len(self.h)
        out = np.zeros(max(n, 1))
        self.L.h_trace(self.h, what.encode(), out)
        return out[:n]

    def bp_trace(self, what):
        n = self.L.h_bp_trace_len(self.h)
        out = np.zeros(max(n, 1))
        self.L.h_bp_trace(self.h, what.encode(), out)
        return out[:n]

    def qp_trace_enable(self, cap):
        self.L.h_qp_trace_enable(self.h, cap)

    def qp_trace(self):
        n = self.L.h_qp_trace_len(self.h)
        ret = np.zeros(max(n, 1), np.int32)
        nf = np.zeros(max(n, 1), np.int32)
        cl = np.zeros(max(n, 1) * self.o.nu, np.int32)
        self.L.h_qp_trace(self.h, ret, nf, cl)
        return ret[:n], nf[:n], cl[: n * self.o.nu].reshape(n, self.o.nu)
