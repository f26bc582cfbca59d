"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference solver core (oracle/_ref, built in
place from /root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The fixtures pin (a) the oracle port and (b) the CUDA path on machines where /root/reference does not exist.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))

import oracle_lib  # noqa: E402
import parity_util as PU  # noqa: E402
from ilqg_b200 import workloads as W  # noqa: E402

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def solve_fixture(problem, ddp, T, params, x0, u0, opts, with_qp=False):
    rec = PU.oracle_record("reference", problem, ddp, T, params, x0, u0, opts, qp_cap=400000 if with_qp else 0)
    out = {k: np.asarray(v) for k, v in rec.items() if k != "qp"}
    if with_qp:
        ret, nfree, cl = rec["qp"]
        out["qp_ret_last"] = ret[-T:]
        out["qp_clamped_last"] = cl[-T:]
        out["qp_ret_hist"] = np.stack(np.unique(ret, return_counts=True))
    out["x0"], out["u0"] = np.asarray(x0), np.asarray(u0)
    return out


def kats():
    """Known-answer tests of the small dense helpers, produced by calling the reference functions directly."""
    lib = C.CDLL(oracle_lib.lib_path("reference", "car", 0))
    rng = np.random.default_rng(12345)
    out = {}
    sym = lambda n: (n * (n + 1)) // 2
    lib.addMulVec.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int]
    lib.addSquareTri.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, _dp]
    lib.addMul2Tri.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp]
    lib.cholesky_tri.argtypes = [_dp, C.c_int, _dp]
    lib.cholesky_tri_inv.argtypes = [_dp, _dp, C.c_int, _dp]
    lib.boxQP.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, C.c_int]
    cases = []
    for case, (nr, nc) in enumerate([(4, 2), (4, 4), (2, 4), (12, 4), (1, 1), (3, 5)]):
        a = rng.standard_normal(nr * nc)
        Bm = rng.standard_normal(sym(nr))
        v = rng.standard_normal(nr)
        base1 = rng.standard_normal(nc)
        r1 = base1.copy(); lib.addMulVec(r1, v, a, nr, nc)
        base2 = rng.standard_normal(sym(nc))
        r2 = base2.copy(); lib.addSquareTri(r2, Bm, a, nr, nc, np.zeros(nr * nc))
        ncc = 3
        c = rng.standard_normal(nr * ncc)
        base3 = rng.standard_normal(nc * ncc)
        r3 = base3.copy(); lib.addMul2Tri(r3, Bm, a, nr, nc, c, nr, ncc, np.zeros(nr * ncc))
        out.update({f"mm{case}_dims": np.array([nr, nc, ncc]), f"mm{case}_a": a, f"mm{case}_B": Bm, f"mm{case}_v": v, f"mm{case}_c": c,
                    f"mm{case}_base1": base1, f"mm{case}_r1": r1, f"mm{case}_base2": base2, f"mm{case}_r2": r2,
                    f"mm{case}_base3": base3, f"mm{case}_r3": r3})
        cases.append(case)
    out["mm_cases"] = np.array(cases)
    # Cholesky / inverse on SPD and on an indefinite matrix
    n_ch = 0
    for n in (1, 2, 3, 4, 6):
        for trial in range(3):
            M = rng.standard_normal((n, n + 2))
            A = M @ M.T + (0.0 if trial < 2 else -3.0) * np.eye(n)
            Ap = np.array([A[r, c] for c in range(n) for r in range(c + 1)])
            U = np.zeros(sym(n)); ok = lib.cholesky_tri(Ap, n, U)
            inv = np.zeros(sym(n))
            if ok:
                lib.cholesky_tri_inv(U, inv, n, np.zeros(n))
            out.update({f"ch{n_ch}_n": np.array([n]), f"ch{n_ch}_A": Ap, f"ch{n_ch}_ok": np.array([ok]), f"ch{n_ch}_U": U, f"ch{n_ch}_inv": inv})
            n_ch += 1
    out["ch_count"] = np.array([n_ch])
    # box QP: random problems incl. tight / inactive bounds and warm starts
    n_qp = 0
    for n in (1, 2, 3, 4):
        for trial in range(12):
            M = rng.standard_normal((n, n + 1))
            H = M @ M.T + (1e-3 if trial % 4 else -0.5) * np.eye(n)
            Hp = np.array([H[r, c] for c in range(n) for r in range(c + 1)])
            g = rng.standard_normal(n) * (3.0 if trial % 3 == 0 else 0.3)
            lo = -np.abs(rng.standard_normal(n)) * (0.1 if trial % 2 else 2.0)
            hi = np.abs(rng.standard_normal(n)) * (0.1 if trial % 5 == 0 else 2.0)
            if trial == 7:
                lo[:] = -np.inf; hi[:] = np.inf
            x = rng.standard_normal(n)
            x_in = x.copy()
            clamped = np.zeros(n, np.int32); nfree = np.zeros(1, np.int32)
            inv = np.zeros(sym(n))
            ret = lib.boxQP(Hp.copy(), g, lo, hi, x, np.zeros(sym(n)), np.zeros(sym(n)), np.zeros(n), np.zeros(n), np.zeros(n),
                            clamped, nfree, inv, n)
            out.update({f"qp{n_qp}_n": np.array([n]), f"qp{n_qp}_H": Hp, f"qp{n_qp}_g": g, f"qp{n_qp}_lo": lo, f"qp{n_qp}_hi": hi,
                        f"qp{n_qp}_x0": x_in, f"qp{n_qp}_x": x, f"qp{n_qp}_ret": np.array([ret]), f"qp{n_qp}_clamped": clamped,
                        f"qp{n_qp}_nfree": nfree, f"qp{n_qp}_inv": inv})
            n_qp += 1
    out["qp_count"] = np.array([n_qp])
    return out


def make_pend():
    """Running equality (hle) + terminal inequality (hfi): four problems per FULL_DDP setting, incl. one that fails."""
    xp, up = W.pend_batch(4)
    for ddp in (0, 1):
        for b in range(4):
            fx = solve_fixture("pend", ddp, W.PEND_T, W.PEND_PARAMS, xp[b], up[b], W.PEND_OPTS, with_qp=True)
            s = oracle_lib.OracleLib("reference", "pend", ddp).solver(W.PEND_T)
            s.set_opts(W.PEND_OPTS); s.set_params(W.PEND_PARAMS); s.init(xp[b], up[b]); s.solve()
            fx["mult_t"] = s.get("mult_t")
            np.savez_compressed(os.path.join(HERE, f"pend_b{b}_ddp{ddp}.npz"), **fx)


def modchol_cases():
    """Inputs for the modified-Cholesky family: positive definite, indefinite, negative definite, singular and zero matrices."""
    rng = np.random.default_rng(2468)
    cases = []
    for n in (1, 2, 3, 4, 6, 12):
        for trial in range(10):
            M = rng.standard_normal((n, n + 1))
            A = M @ M.T
            if trial % 5 == 1:
                A -= (0.5 + trial) * np.eye(n)                       # indefinite
            elif trial % 5 == 2:
                A = -A - 0.1 * np.eye(n)                             # negative definite
            elif trial % 5 == 3:
                A[:, -1] = A[:, 0]; A[-1, :] = A[0, :]; A[-1, -1] = A[0, 0]   # singular (repeated row / column)
            elif trial == 4:
                A = np.zeros((n, n))
            elif trial == 9:
                A = np.diag(rng.uniform(-1, 1, n))                   # diagonal with mixed signs
            cases.append((n, np.array([A[r, c] for c in range(n) for r in range(c + 1)]), rng.standard_normal(n)))
    return cases


def run_modchol(lib, n, Ap, b):
    sym = (n * (n + 1)) // 2
    L = Ap.copy(); E = np.zeros(n); P = np.zeros(n, np.int32); g = np.zeros(n)
    ret = lib.mod_chol(L, n, E, P, g)
    inv = np.zeros(sym); lib.mod_chol_inv(L, P, inv, n, np.zeros(n))
    H = np.zeros(sym); lib.perm_tri_square(L, H, P, n)
    x = np.zeros(n); lib.mod_chol_solve(L, P, np.ascontiguousarray(b), x, n, np.zeros(n))
    return dict(L=L, E=E, P=P, ret=np.array([ret]), inv=inv, H=H, x=x)


def modchol_lib(path):
    lib = C.CDLL(path)
    lib.mod_chol.restype = C.c_double
    lib.mod_chol.argtypes = [_dp, C.c_int, _dp, _ip, _dp]
    lib.mod_chol_inv.argtypes = [_dp, _ip, _dp, C.c_int, _dp]
    lib.perm_tri_square.argtypes = [_dp, _dp, _ip, C.c_int]
    lib.mod_chol_solve.argtypes = [_dp, _ip, _dp, _dp, C.c_int, _dp]
    return lib


def make_modchol():
    """Known answers of mod_chol / mod_chol_inv / perm_tri_square / mod_chol_solve (cholesky.c:129-356), from the reference."""
    lib = modchol_lib(oracle_lib.lib_path("reference", "car", 0))
    out = {}
    cases = modchol_cases()
    for i, (n, Ap, b) in enumerate(cases):
        r = run_modchol(lib, n, Ap, b)
        out.update({f"c{i}_n": np.array([n]), f"c{i}_A": Ap, f"c{i}_b": b, **{f"c{i}_{k}": v for k, v in r.items()}})
    out["count"] = np.array([len(cases)])
    np.savez_compressed(os.path.join(HERE, "modchol_kats.npz"), **out)


def main():
    assert oracle_lib.available("reference", "car", 0), "build oracle/_ref first (make -C oracle ref)"
    x0, u0 = W.car_single()
    for ddp in (0, 1):
        np.savez_compressed(os.path.join(HERE, f"car_single_ddp{ddp}.npz"),
                            **solve_fixture("car", ddp, 500, W.CAR_PARAMS, x0, u0, {"max_iter": 200}, with_qp=True))
    xb, ub = W.car_batch(8, T=100, seed=5)
    for b in range(8):
        np.savez_compressed(os.path.join(HERE, f"car_T100_b{b}.npz"),
                            **solve_fixture("car", 0, 100, W.CAR_PARAMS, xb[b], ub[b], {"max_iter": 30}))
    for n in (2, 3, 5, 500):
        params, bx0, bu0, opts = W.brachi(n)
        for ddp in (0, 1):
            np.savez_compressed(os.path.join(HERE, f"brachi_n{n}_ddp{ddp}.npz"),
                                **solve_fixture("brachi", ddp, n, params, bx0, bu0, opts))
    xq, uq = W.quad_batch(2, T=300)
    for b in range(2):
        for ddp in (0, 1):
            np.savez_compressed(os.path.join(HERE, f"quad_T300_b{b}_ddp{ddp}.npz"),
                                **solve_fixture("quad", ddp, 300, W.QUAD_PARAMS, xq[b], uq[b], {"max_iter": 25}))
    xs, us = W.car_single()
    for ddp in (0, 1):
        np.savez_compressed(os.path.join(HERE, f"carhx_ddp{ddp}.npz"),
                            **solve_fixture("carhx", ddp, 500, W.CARHX_PARAMS, xs, us, {"max_iter": 60}, with_qp=True))
    params, bx0, bu0, opts = W.brachi_hli(500)
    for ddp in (0, 1):
        fx = solve_fixture("brachi_hli", ddp, 500, params, bx0, bu0, opts)
        s = oracle_lib.OracleLib("reference", "brachi_hli", ddp).solver(500)
        s.set_opts(opts); s.set_params(params); s.init(bx0, bu0); s.solve()
        fx["mult_t"] = s.get("mult_t")
        np.savez_compressed(os.path.join(HERE, f"brachi_hli_ddp{ddp}.npz"), **fx)
    make_pend()
    make_modchol()
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **kats())
    # bit patterns of the deterministic math layer on a fixed grid
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libdmcheck.so"))
    lib.dmc_vec.argtypes = [C.c_int, _dp, _dp, C.c_long]
    xs = np.concatenate([np.linspace(-7, 7, 2001), np.linspace(-1e4, 1e4, 501), [0.0, -0.0, 1e-300, 1.5e6, 1.7e6]])
    xa = np.concatenate([np.linspace(-1, 1, 2001), [0.5, -0.5, 0.975, 1e-9, 1.0000001]])
    ys = {}
    for i, (nm, grid) in enumerate([("sin", xs), ("cos", xs), ("asin", xa), ("acos", xa)]):
        y = np.zeros_like(grid); lib.dmc_vec(i, np.ascontiguousarray(grid), y, grid.size); ys[nm] = y
    np.savez_compressed(os.path.join(HERE, "dm_math.npz"), xs=xs, xa=xa, **ys)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pend":      # only the fixtures added in round 2 (the others stay byte-identical)
        make_pend()
    elif len(sys.argv) > 1 and sys.argv[1] == "modchol":
        make_modchol()
    else:
        main()
