"""The modified-Cholesky family (cholesky.c:129-356): known answers recorded from the reference's own
mod_chol / mod_chol_inv / perm_tri_square / mod_chol_solve (tests/golden/modchol_kats.npz, make_golden.py) against the product's
device unit csrc/mod_chol.cuh -- compiled for the host (no GPU needed) and run in a kernel through ilqgb_mod_chol (GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

import ilqg_b200
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "modchol_kats.npz")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
KEYS = ("L", "E", "P", "ret", "inv", "H", "x")


def cases():
    g = np.load(GOLD)
    return [{k: g[f"c{i}_{k}"] for k in ("n", "A", "b") + KEYS} for i in range(int(g["count"][0]))]


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def run_functions(fn, n, Ap, b):
    sym = (n * (n + 1)) // 2
    L = Ap.copy(); E = np.zeros(n); P = np.zeros(n, np.int32)
    ret = fn["mod_chol"](L, n, E, P, np.zeros(n))
    inv = np.zeros(sym); fn["inv"](L, P, inv, n, np.zeros(n))
    H = np.zeros(sym); fn["sq"](L, H, P, n)
    x = np.zeros(n); fn["solve"](L, P, np.ascontiguousarray(b), x, n, np.zeros(n))
    return dict(L=L, E=E, P=P, ret=np.array([ret]), inv=inv, H=H, x=x)


def bind(lib, names):
    fn = {k: getattr(lib, v) for k, v in names.items()}
    fn["mod_chol"].restype = C.c_double
    fn["mod_chol"].argtypes = [_dp, C.c_int, _dp, _ip, _dp]
    fn["inv"].argtypes = [_dp, _ip, _dp, C.c_int, _dp]
    fn["sq"].argtypes = [_dp, _dp, _ip, C.c_int]
    fn["solve"].argtypes = [_dp, _ip, _dp, _dp, C.c_int, _dp]
    return fn


def test_fixture_covers_every_branch():
    cs = cases()
    assert len(cs) == 60 and {int(c["n"][0]) for c in cs} == {1, 2, 3, 4, 6, 12}
    shifted = [c for c in cs if c["ret"][0] > 0]
    assert 20 <= len(shifted) <= 50                                       # phase two taken and not taken
    assert any((np.diff(c["P"]) < 0).any() for c in cs)                   # pivoting happened
    # where no shift was needed the factor is a plain pivoted Cholesky factor: P L'L P' gives A back.  (Not for every unshifted
    # case: phase two can end with a zero shift and then goes through the final 2 x 2 block, which reads element (0, n-1) for
    # (n-2, n-1) -- a quirk of the reference that the restatement keeps, cholesky.c:281-297.)
    plain = [c for c in cs if int(c["n"][0]) > 1 and c["ret"][0] == 0 and np.isfinite(c["L"]).all()]
    assert sum(np.allclose(c["H"], c["A"], rtol=1e-10, atol=1e-12) for c in plain) >= 10


def test_reference_reproduces_the_fixture():
    if not oracle_lib.available("reference", "car", 0):
        pytest.skip("reference core not built on this machine")
    fn = bind(C.CDLL(oracle_lib.lib_path("reference", "car", 0)), {"mod_chol": "mod_chol", "inv": "mod_chol_inv", "sq": "perm_tri_square", "solve": "mod_chol_solve"})
    for i, c in enumerate(cases()):
        r = run_functions(fn, int(c["n"][0]), c["A"], c["b"])
        for k in KEYS:
            assert same(r[k], c[k]), (i, k)


def test_device_unit_compiled_for_the_host_matches_the_reference():
    path = os.path.join(ROOT, "oracle", "_build", "libmodchol_host.so")
    assert os.path.exists(path), "make -C oracle port"
    fn = bind(C.CDLL(path), {"mod_chol": "mch_mod_chol", "inv": "mch_inv", "sq": "mch_perm_tri_square", "solve": "mch_solve"})
    for i, c in enumerate(cases()):
        r = run_functions(fn, int(c["n"][0]), c["A"], c["b"])
        for k in KEYS:
            assert same(r[k], c[k]), (i, k)


@pytest.mark.gpu
def test_device_unit_on_the_gpu_matches_the_reference():
    lib = ilqg_b200.Library("car", 0).lib
    lib.ilqgb_mod_chol.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp]
    cs = cases()
    for n in sorted({int(c["n"][0]) for c in cs}):
        grp = [c for c in cs if int(c["n"][0]) == n]
        sym = (n * (n + 1)) // 2
        A = np.ascontiguousarray(np.stack([c["A"] for c in grp])); b = np.ascontiguousarray(np.stack([c["b"] for c in grp]))
        fac, inv, H = (np.zeros((len(grp), sym)) for _ in range(3))
        E, x = np.zeros((len(grp), n)), np.zeros((len(grp), n))
        P = np.zeros((len(grp), n), np.int32); ret = np.zeros(len(grp))
        assert lib.ilqgb_mod_chol(0, n, len(grp), A, b, fac, E, P, ret, inv, H, x) == 0
        for i, c in enumerate(grp):
            got = dict(L=fac[i], E=E[i], P=P[i], ret=ret[i:i + 1], inv=inv[i], H=H[i], x=x[i])
            for k in KEYS:
                assert same(got[k], c[k]), (n, i, k)
