"""Deterministic math layer (csrc/dm_math.h): bit patterns pinned by a golden grid, accuracy against libm."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _lib():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libdmcheck.so"))
    lib.dmc_vec.argtypes = [C.c_int, _dp, _dp, C.c_long]
    return lib


def _eval(lib, which, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    lib.dmc_vec(which, x, y, x.size)
    return y


def test_bits_match_golden_grid():
    g = np.load(os.path.join(ROOT, "tests", "golden", "dm_math.npz"))
    lib = _lib()
    for i, (nm, grid) in enumerate([("sin", g["xs"]), ("cos", g["xs"]), ("asin", g["xa"]), ("acos", g["xa"])]):
        assert np.array_equal(_eval(lib, i, grid), g[nm], equal_nan=True), nm


def test_within_one_ulp_of_libm():
    lib = _lib()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 200000), rng.uniform(-10, 10, 200000), rng.uniform(-1e3, 1e3, 200000), rng.uniform(-1.5e6, 1.5e6, 200000)])
    for i, f in ((0, np.sin), (1, np.cos)):
        y, r = _eval(lib, i, x), f(x)
        assert np.max(np.abs(y - r) / np.spacing(np.abs(r))) <= 1.0
    a = rng.uniform(-1, 1, 400000)
    y, r = _eval(lib, 2, a), np.arcsin(a)
    assert np.max(np.abs(y - r) / np.spacing(np.abs(r))) <= 1.0
    y, r = _eval(lib, 3, a), np.arccos(a)
    assert np.max(np.abs(y - r) / np.spacing(np.abs(r))) <= 1.0


def test_out_of_range_is_nan():
    lib = _lib()
    assert np.isnan(_eval(lib, 0, [2e6, np.inf, np.nan])).all()
    assert np.isnan(_eval(lib, 2, [1.0000001, -2.0])).all()
    assert _eval(lib, 2, [1.0])[0] == np.pi / 2
