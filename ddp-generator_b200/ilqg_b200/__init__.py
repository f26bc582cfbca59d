"""Host-side Python mirror of the C ABI in include/ilqg_b200.h (ctypes; no torch types in any signature).

Usage mirrors the reference's mex call `[success, x, u, cost] = iLQG<Name>(x0, u0, p, Op)` (iLQG_mex.c:19-33) for a
whole batch:

    s = BatchSolver("car", full_ddp=0, batch=B, n_hor=T)
    s.set_options({"max_iter": 50}); s.set_params(params)
    out = s.solve(x0, u0)        # dict: success, x, u, cost, iterations, n_linesearch

The shared library is the product; if it is missing or no GPU is usable, construction raises -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("ILQG_LIB_DIR") or os.path.join(os.path.dirname(_HERE), "lib")
TRACE, TIMING = 1, 2
KERNEL_CLASSES = ("derivs", "backpass", "linesearch", "post")
_SCALAR_FIELDS = {"cost", "new_cost", "dcost", "expected", "lambda", "dlambda", "g_norm", "dV0", "dV1", "w_pen_l", "w_pen_f",
                  "iterations", "result", "status", "n_linesearch", "n_backpass", "n_derivs", "n_rollouts", "n_tails", "cur", "bp_split"}


def lib_path(problem, full_ddp, lib_dir=None):
    return os.path.join(lib_dir or LIB_DIR, f"libilqg_b200_{problem}_ddp{int(full_ddp)}.so")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Library:
    """One loaded libilqg_b200_<problem>_ddp<d>.so with typed entry points."""

    _cache = {}

    def __new__(cls, problem, full_ddp, lib_dir=None):
        key = (problem, int(full_ddp), lib_dir)
        if key not in cls._cache:
            self = super().__new__(cls)
            self._load(problem, full_ddp, lib_dir)
            cls._cache[key] = self
        return cls._cache[key]

    def _load(self, problem, full_ddp, lib_dir=None):
        path = lib_path(problem, full_ddp, lib_dir)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found: build it with `make -C ddp-generator_b200` (no CPU fallback exists)")
        L = self.lib = C.CDLL(path)
        self.path = path
        vp, cp, ci, dp = C.c_void_p, C.c_char_p, C.c_int, C.c_void_p
        L.ilqgb_problem_name.restype = cp
        L.ilqgb_param_name.restype = cp
        L.ilqgb_param_name.argtypes = [ci]
        L.ilqgb_param_size.argtypes = [ci]
        L.ilqgb_create.restype = vp
        L.ilqgb_create.argtypes = [ci, ci, ci, ci, vp]
        L.ilqgb_create_multi.restype = vp
        L.ilqgb_create_multi.argtypes = [ci, vp, ci, ci, ci, vp]
        L.ilqgb_devices.argtypes = [vp]
        L.ilqgb_destroy.argtypes = [vp]
        L.ilqgb_last_error.restype = cp
        L.ilqgb_last_error.argtypes = [vp]
        L.ilqgb_standard_parameters.argtypes = [vp]
        L.ilqgb_set_opt.restype = cp
        L.ilqgb_set_opt.argtypes = [vp, cp, dp, ci]
        L.ilqgb_set_param.argtypes = [vp, ci, dp, ci]
        L.ilqgb_set_param_batch.argtypes = [vp, ci, dp, ci]
        L.ilqgb_validate_opt.restype = cp
        L.ilqgb_validate_opt.argtypes = [cp, dp, ci]
        L.ilqgb_upload.argtypes = [vp, dp, dp]
        for f in ("ilqgb_start", "ilqgb_finish", "ilqgb_solve", "ilqgb_sync", "ilqgb_active", "ilqgb_phase_derivs",
                  "ilqgb_phase_backpass", "ilqgb_phase_linesearch"):
            getattr(L, f).argtypes = [vp]
        L.ilqgb_iterate.argtypes = [vp, ci]
        L.ilqgb_download.argtypes = [vp, dp, dp, dp, dp, dp, dp]
        L.ilqgb_solve_host.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.ilqgb_get.restype = C.c_long
        L.ilqgb_get.argtypes = [vp, cp, dp]
        L.ilqgb_get_int.restype = C.c_long
        L.ilqgb_get_int.argtypes = [vp, cp, dp]
        L.ilqgb_timing.argtypes = [vp, dp, dp, ci]
        L.ilqgb_launch_count.restype = C.c_long
        L.ilqgb_launch_count.argtypes = [vp]
        L.ilqgb_chunks.argtypes = [vp]
        L.ilqgb_set_tuning.argtypes = [vp, cp, ci]
        self.problem = L.ilqgb_problem_name().decode()
        self.nx, self.nu = L.ilqgb_nx(), L.ilqgb_nu()
        self.full_ddp = L.ilqgb_full_ddp()
        self.param_names = [L.ilqgb_param_name(i).decode() for i in range(L.ilqgb_n_params())]
        self.param_sizes = [L.ilqgb_param_size(i) for i in range(L.ilqgb_n_params())]
        self.deriv_doubles_per_step = L.ilqgb_deriv_doubles_per_step()

    def validate_option(self, name, value):
        """setOptParam's check without a handle: None or the reference's message."""
        v = np.ascontiguousarray(np.atleast_1d(np.asarray(value, dtype=np.float64)))
        err = self.lib.ilqgb_validate_opt(name.encode(), _ptr(v), v.size)
        return err.decode() if err else None

    def device_count(self):
        return self.lib.ilqgb_device_count()


class BatchSolver:
    def __init__(self, problem, full_ddp=0, batch=1, n_hor=1, device=0, flags=0, stream=None, chunks=0, devices=None, lib_dir=None):
        """`devices`: list of GPU indices (or a count) to shard the batch over in this one process (ilqgb_create_multi);
        `chunks` then counts chunks per device.  Default: the single GPU `device`.  `lib_dir`: where the problem's library was
        built (python -m ilqg_gen.make --out DIR puts it in DIR/lib); default: the package's lib/."""
        self.L = Library(problem, full_ddp, lib_dir)
        self.lib = self.L.lib
        self.B, self.T = int(batch), int(n_hor)
        self.nx, self.nu = self.L.nx, self.L.nu
        self.max_iter = 20
        fl = int(flags) | ((int(chunks) & 0xff) << 8)
        st = C.c_void_p(stream) if stream else None
        if devices is None:
            self.h = self.lib.ilqgb_create(int(device), self.B, self.T, fl, st)
        else:
            devs = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
            arr = (C.c_int * len(devs))(*devs)
            self.h = self.lib.ilqgb_create_multi(len(devs), arr, self.B, self.T, fl, st)
        if not self.h:
            raise RuntimeError("ilqgb_create failed: " + self.lib.ilqgb_last_error(None).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.ilqgb_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(self.lib.ilqgb_last_error(self.h).decode())
        return rc

    # options / parameters ------------------------------------------------------------------------------------------
    def set_option(self, name, value):
        """Returns None or the reference's error message (setOptParam semantics)."""
        v = np.ascontiguousarray(np.atleast_1d(np.asarray(value, dtype=np.float64)))
        err = self.lib.ilqgb_set_opt(self.h, name.encode(), _ptr(v), v.size)
        if err is None and name == "max_iter":
            self.max_iter = int(v[0])
        return err.decode() if err else None

    def set_options(self, opts):
        for k, v in opts.items():
            err = self.set_option(k, v)
            if err:
                raise ValueError(f"Error setting optimization parameter '{k}': {err}.")

    def set_params(self, params):
        for i, name in enumerate(self.L.param_names):
            if name not in params:
                raise KeyError(f"Parameter name '{name}' is not member of parameters struct.")
            v = np.ascontiguousarray(np.asarray(params[name], dtype=np.float64).ravel())
            self._chk(self.lib.ilqgb_set_param(self.h, i, _ptr(v), v.size))

    def set_params_batch(self, params):
        """Per-problem parameter sets: {name: array [batch, size]} for the names that differ between problems."""
        for name, val in params.items():
            i = self.L.param_names.index(name)
            v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).reshape(self.B, -1))
            self._chk(self.lib.ilqgb_set_param_batch(self.h, i, _ptr(v), v.shape[1]))

    # data ----------------------------------------------------------------------------------------------------------------
    def upload(self, x0, u0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        if x0.shape != (self.B, self.nx):
            raise ValueError(f"wrong number of elements in x0 ({self.B}x{self.nx} expected)")
        if u0.shape != (self.B, self.T, self.nu):
            raise ValueError(f"wrong number of elements in u_nom ({self.B}x{self.T}x{self.nu} expected)")
        self._chk(self.lib.ilqgb_upload(self.h, _ptr(x0), _ptr(u0)))
        self._keep = (x0, u0)

    def upload_ptr(self, x0_ptr, u0_ptr):
        self._chk(self.lib.ilqgb_upload(self.h, C.c_void_p(x0_ptr), C.c_void_p(u0_ptr)))

    def download(self, want_traj=True):
        x = np.empty((self.B, self.T + 1, self.nx)) if want_traj else None
        u = np.empty((self.B, self.T, self.nu)) if want_traj else None
        cost = np.empty(self.B)
        it = np.empty(self.B, np.int32)
        res = np.empty(self.B, np.int32)
        nls = np.empty(self.B, np.int32)
        self._chk(self.lib.ilqgb_download(self.h, _ptr(x), _ptr(u), _ptr(cost), _ptr(it), _ptr(res), _ptr(nls)))
        return dict(success=res, x=x, u=u, cost=cost, iterations=it, n_linesearch=nls)

    def download_ptr(self, x_ptr, u_ptr, cost_ptr, it_ptr, res_ptr, nls_ptr):
        self._chk(self.lib.ilqgb_download(self.h, *[C.c_void_p(p) if p else None for p in (x_ptr, u_ptr, cost_ptr, it_ptr, res_ptr, nls_ptr)]))

    # solve -------------------------------------------------------------------------------------------------------------
    def start(self):
        self._chk(self.lib.ilqgb_start(self.h))

    def iterate(self, n):
        return self._chk(self.lib.ilqgb_iterate(self.h, int(n)))

    def finish(self):
        self._chk(self.lib.ilqgb_finish(self.h))

    def run(self):
        self._chk(self.lib.ilqgb_solve(self.h))

    def sync(self):
        self._chk(self.lib.ilqgb_sync(self.h))

    def active(self):
        return self._chk(self.lib.ilqgb_active(self.h))

    def solve(self, x0, u0, want_traj=True):
        self.upload(x0, u0)
        self.run()
        return self.download(want_traj)

    def solve_host_ptr(self, x0_ptr, u0_ptr, x_ptr, u_ptr, cost_ptr, it_ptr, res_ptr, nls_ptr):
        """Pipelined upload + solve + download on raw host pointers (pinned memory for full overlap)."""
        self._chk(self.lib.ilqgb_solve_host(self.h, *[C.c_void_p(p) if p else None for p in (x0_ptr, u0_ptr, x_ptr, u_ptr, cost_ptr, it_ptr, res_ptr, nls_ptr)]))

    def solve_host(self, x0, u0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        x = np.empty((self.B, self.T + 1, self.nx)); u = np.empty((self.B, self.T, self.nu)); cost = np.empty(self.B)
        it = np.empty(self.B, np.int32); res = np.empty(self.B, np.int32); nls = np.empty(self.B, np.int32)
        self._chk(self.lib.ilqgb_solve_host(self.h, _ptr(x0), _ptr(u0), _ptr(x), _ptr(u), _ptr(cost), _ptr(it), _ptr(res), _ptr(nls)))
        return dict(success=res, x=x, u=u, cost=cost, iterations=it, n_linesearch=nls)

    def set_tuning(self, name, value):
        """Run-time tuning knobs (ilqgb_set_tuning): ls_tail_from, bp_latency, bp_split, cw_lpp, pass_index."""
        self._chk(self.lib.ilqgb_set_tuning(self.h, name.encode(), int(value)))

    def phase(self, which):
        self._chk(getattr(self.lib, f"ilqgb_phase_{which}")(self.h))

    # read-back ---------------------------------------------------------------------------------------------------------
    def get(self, field):
        T, nx, nu, B = self.T, self.nx, self.nu, self.B
        nq = nx * (nx + 1) // 2
        shapes = {"x": (B, T + 1, nx), "u": (B, T, nu), "l": (B, T, nu), "L": (B, T, nu * nx), "fd": (B, nx + nq)}
        n_max = B * max((T + 1) * max(nx * nu, self.L.deriv_doubles_per_step, nx, 4), self.max_iter + 1) + 64
        buf = np.empty(n_max)
        n = self._chk(self.lib.ilqgb_get(self.h, field.encode(), _ptr(buf)))
        out = buf[:n].copy()
        if field in shapes:
            return out.reshape(shapes[field])
        return out if field in _SCALAR_FIELDS else out.reshape(B, -1)

    def get_int(self, field):
        n_max = self.B * max(self.T, self.max_iter + 1) + 64
        buf = np.empty(n_max, np.int32)
        n = self._chk(self.lib.ilqgb_get_int(self.h, field.encode(), _ptr(buf)))
        out = buf[:n].copy()
        return out if field in _SCALAR_FIELDS else out.reshape(self.B, -1)

    def chunks(self):
        return int(self.lib.ilqgb_chunks(self.h))

    def devices(self):
        return int(self.lib.ilqgb_devices(self.h))

    def launch_count(self):
        return int(self.lib.ilqgb_launch_count(self.h))

    def timing(self, reset=True):
        ms = np.zeros(4)
        n = np.zeros(4, np.int64)
        self._chk(self.lib.ilqgb_timing(self.h, _ptr(ms), _ptr(n), int(reset)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}
