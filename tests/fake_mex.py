"""Drive a mexFunction gateway from Python through oracle/mex_stub/fake_mex.c (TEST INFRASTRUCTURE).

`libfakemex.so` is loaded RTLD_GLOBAL first so that a gateway library -- the reference's iLQG_mex.c built in
oracle/_ref/libmexref_*.so, or this repo's ddp-generator_b200/mex/iLQG_mex_b200.c built in oracle/_build/libmexb200_*.so --
finds the mx*/mex* symbols it leaves undefined, exactly as a real mex file does inside MATLAB / Octave.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKEMEX = os.path.join(ROOT, "oracle", "_build", "libfakemex.so")


def gateway_path(kind, problem, full_ddp=0):
    """kind: 'reference' (iLQG_mex.c + reference core, CPU) or 'b200' (this repo's gateway over the GPU library)."""
    if kind == "reference":
        return os.path.join(ROOT, "oracle", "_ref", f"libmexref_{problem}_ddp{int(full_ddp)}.so")
    return os.path.join(ROOT, "oracle", "_build", f"libmexb200_{problem}_ddp{int(full_ddp)}.so")


class MexError(Exception):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.ident, self.msg = ident, msg


_rt = None


def runtime():
    global _rt
    if _rt is None:
        L = C.CDLL(FAKEMEX, mode=C.RTLD_GLOBAL)
        L.fm_new_double.restype = C.c_void_p
        L.fm_new_double.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.c_void_p]
        L.fm_new_struct.restype = C.c_void_p
        L.fm_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.fm_mark_sparse.argtypes = [C.c_void_p, C.c_int]
        L.fm_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
        L.fm_error_id.restype = C.c_char_p
        L.fm_error_msg.restype = C.c_char_p
        L.fm_printed.restype = C.c_char_p
        L.fm_live_allocs.restype = C.c_long
        L.mxDestroyArray.argtypes = [C.c_void_p]
        L.mxGetNumberOfDimensions.restype = C.c_size_t
        L.mxGetNumberOfDimensions.argtypes = [C.c_void_p]
        L.mxGetDimensions.restype = C.POINTER(C.c_size_t)
        L.mxGetDimensions.argtypes = [C.c_void_p]
        L.mxGetPr.restype = C.POINTER(C.c_double)
        L.mxGetPr.argtypes = [C.c_void_p]
        _rt = L
    return _rt


class Sparse:
    """Marks an input as a (fake) sparse matrix, to exercise the gateway's type checks."""

    def __init__(self, a):
        self.a = a


def to_mx(v):
    L = runtime()
    if isinstance(v, dict):
        s = L.fm_new_struct()
        for k, val in v.items():
            L.fm_set_field(s, k.encode(), to_mx(val))
        return s
    sparse = isinstance(v, Sparse)
    a = np.asarray(v.a if sparse else v, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)            # a vector is a column, as x0 = [..]' in the reference's test scripts
    dims = (C.c_size_t * a.ndim)(*a.shape)
    flat = np.ascontiguousarray(a.ravel(order="F"))
    m = L.fm_new_double(a.ndim, dims, flat.ctypes.data_as(C.c_void_p))
    if sparse:
        L.fm_mark_sparse(m, 1)
    return m


def from_mx(p):
    L = runtime()
    nd = L.mxGetNumberOfDimensions(p)
    dims = [L.mxGetDimensions(p)[i] for i in range(nd)]
    n = int(np.prod(dims))
    pr = L.mxGetPr(p)
    return np.array(pr[:n], dtype=np.float64).reshape(dims, order="F")


class Gateway:
    def __init__(self, path):
        self.rt = runtime()
        self.lib = C.CDLL(path)
        self.fn = C.cast(self.lib.mexFunction, C.c_void_p)

    def __call__(self, *args, nlhs=4):
        """Returns the nlhs outputs as numpy arrays; raises MexError on mexErrMsgIdAndTxt."""
        L = self.rt
        ins = [to_mx(a) for a in args]
        prhs = (C.c_void_p * max(len(ins), 1))(*ins)
        plhs = (C.c_void_p * max(nlhs, 1))()
        rc = L.fm_call(self.fn, nlhs, plhs, len(ins), prhs)
        self.printed = L.fm_printed().decode(errors="replace")
        self.live_allocs = L.fm_live_allocs()
        try:
            if rc:
                raise MexError(L.fm_error_id().decode(), L.fm_error_msg().decode())
            return [from_mx(plhs[i]) for i in range(nlhs)]
        finally:
            for a in ins:
                L.mxDestroyArray(a)
            for i in range(nlhs):
                if plhs[i]:
                    L.mxDestroyArray(plhs[i])
