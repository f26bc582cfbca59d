"""Pins the oracle: the C port (oracle/port_*.c) -- and the compiled reference when it is present -- must reproduce
the golden fixtures that tests/golden/make_golden.py recorded from the UNMODIFIED reference solver core, bit for bit:
full solves (iterations, lambda / alpha / cost sequences, trajectories, multipliers, box-QP active sets) and unit-level
known answers of the dense helpers."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import oracle_lib
import parity_util as PU
from ilqg_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
KEYS = ("result", "iterations", "n_ls", "n_bp", "cost", "cost0", "lambda", "g_norm", "w_pen_l", "w_pen_f", "dV0", "dV1",
        "tr_alpha", "tr_lambda", "tr_newcost", "x", "u", "l", "L", "mult_f")


def kinds(problem, ddp):
    k = PU.oracle_kinds(problem, ddp)
    assert "port" in k, "oracle port not built: run `make -C oracle port` (or __graft_entry__.build())"
    return k


def check_solve(fixture, problem, ddp, T, params, opts, with_qp=False):
    g = np.load(os.path.join(GOLD, fixture))
    for kind in kinds(problem, ddp):
        rec = PU.oracle_record(kind, problem, ddp, T, params, g["x0"], g["u0"], opts, qp_cap=400000 if with_qp else 0)
        for k in KEYS:
            assert np.array_equal(np.asarray(rec[k]), g[k]), f"{fixture}: {k} differs for {kind}"
        if with_qp:
            ret, _, cl = rec["qp"]
            assert np.array_equal(ret[-T:], g["qp_ret_last"]) and np.array_equal(cl[-T:], g["qp_clamped_last"])
            assert np.array_equal(np.stack(np.unique(ret, return_counts=True)), g["qp_ret_hist"])


@pytest.mark.parametrize("ddp", [0, 1])
def test_car_single(ddp):
    check_solve(f"car_single_ddp{ddp}.npz", "car", ddp, 500, W.CAR_PARAMS, {"max_iter": 200}, with_qp=True)


@pytest.mark.parametrize("b", range(8))
def test_car_short(b):
    check_solve(f"car_T100_b{b}.npz", "car", 0, 100, W.CAR_PARAMS, {"max_iter": 30})


@pytest.mark.parametrize("n", [2, 3, 5, 500])
@pytest.mark.parametrize("ddp", [0, 1])
def test_brachistochrone(n, ddp):
    params, _, _, opts = W.brachi(n)
    check_solve(f"brachi_n{n}_ddp{ddp}.npz", "brachi", ddp, n, params, opts)


@pytest.mark.parametrize("b", [0, 1])
@pytest.mark.parametrize("ddp", [0, 1])
def test_quadrotor(b, ddp):
    check_solve(f"quad_T300_b{b}_ddp{ddp}.npz", "quad", ddp, 300, W.QUAD_PARAMS, {"max_iter": 25})


@pytest.mark.parametrize("ddp", [0, 1])
def test_car_state_dependent_limits(ddp):
    check_solve(f"carhx_ddp{ddp}.npz", "carhx", ddp, 500, W.CARHX_PARAMS, {"max_iter": 60}, with_qp=True)


@pytest.mark.parametrize("ddp", [0, 1])
def test_brachistochrone_running_inequality(ddp):
    params, _, _, opts = W.brachi_hli(500)
    check_solve(f"brachi_hli_ddp{ddp}.npz", "brachi_hli", ddp, 500, params, opts)
    g = np.load(os.path.join(GOLD, f"brachi_hli_ddp{ddp}.npz"))
    assert (g["x"][:, 0] - params["ymin"]).min() > -1e-5      # the path respects the running bound y >= ymin[k]


@pytest.mark.parametrize("b", range(4))
@pytest.mark.parametrize("ddp", [0, 1])
def test_pendulum_running_equality_and_terminal_inequality(b, ddp):
    """hle (running equality, iLQG_func.tem:427-453) and hfi (terminal inequality, iLQG_func.tem:470-509): the two constraint
    kinds no reference example uses.  Fixture b=1/ddp=1 is a failed solve (result 0)."""
    check_solve(f"pend_b{b}_ddp{ddp}.npz", "pend", ddp, W.PEND_T, W.PEND_PARAMS, W.PEND_OPTS, with_qp=True)
    g = np.load(os.path.join(GOLD, f"pend_b{b}_ddp{ddp}.npz"))
    for kind in kinds("pend", ddp):
        s = oracle_lib.OracleLib(kind, "pend", ddp).solver(W.PEND_T)
        s.set_opts(W.PEND_OPTS); s.set_params(W.PEND_PARAMS); s.init(g["x0"], g["u0"]); s.solve()
        assert np.array_equal(s.get("mult_t"), g["mult_t"]), kind      # mu_le[k], last_hle[k]
        s.close()
    if int(g["result"]) == 1:       # the fixture itself: both constraints are active and satisfied at the solution
        assert np.abs(g["u"][:, 0] - g["u"][:, 1]).max() < 1e-6
        assert abs(g["x"][-1, 0] - W.PEND_PARAMS["thmin"][0]) < 1e-3 and g["mult_f"][0] > 1.0
        assert np.abs(g["mult_t"][:, 0]).max() > 1e-3


def test_brachistochrone_reaches_cycloid_time():
    """Sanity of the fixture itself: n=500 converges towards the analytic cycloid time pi*sqrt(2/g)."""
    g = np.load(os.path.join(GOLD, "brachi_n500_ddp0.npz"))
    assert abs(float(g["cost"]) - np.pi * np.sqrt(2 / 9.81)) < 5e-4
    assert abs(g["x"][-1, 0] + 4.0) < 1e-6


def _lib(kind):
    lib = C.CDLL(oracle_lib.lib_path(kind, "car", 0))
    lib.addMulVec.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int]
    lib.addSquareTri.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, _dp]
    lib.addMul2Tri.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp]
    lib.cholesky_tri.argtypes = [_dp, C.c_int, _dp]
    lib.cholesky_tri_inv.argtypes = [_dp, _dp, C.c_int, _dp]
    lib.boxQP.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, C.c_int]
    return lib


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_dense_helper_kats(kind):
    if not oracle_lib.available(kind, "car", 0):
        pytest.skip(f"{kind} library not built here")
    lib, g = _lib(kind), np.load(os.path.join(GOLD, "kats.npz"))
    sym = lambda n: (n * (n + 1)) // 2
    for case in g["mm_cases"]:
        nr, nc, ncc = (int(v) for v in g[f"mm{case}_dims"])
        r1 = g[f"mm{case}_base1"].copy(); lib.addMulVec(r1, g[f"mm{case}_v"], g[f"mm{case}_a"], nr, nc)
        r2 = g[f"mm{case}_base2"].copy(); lib.addSquareTri(r2, g[f"mm{case}_B"], g[f"mm{case}_a"], nr, nc, np.zeros(nr * nc))
        r3 = g[f"mm{case}_base3"].copy(); lib.addMul2Tri(r3, g[f"mm{case}_B"], g[f"mm{case}_a"], nr, nc, g[f"mm{case}_c"], nr, ncc, np.zeros(nr * ncc))
        assert np.array_equal(r1, g[f"mm{case}_r1"]) and np.array_equal(r2, g[f"mm{case}_r2"]) and np.array_equal(r3, g[f"mm{case}_r3"])
    for i in range(int(g["ch_count"][0])):
        n = int(g[f"ch{i}_n"][0])
        U = np.zeros(sym(n)); ok = lib.cholesky_tri(g[f"ch{i}_A"], n, U)
        assert ok == int(g[f"ch{i}_ok"][0])
        if ok:
            inv = np.zeros(sym(n)); lib.cholesky_tri_inv(U, inv, n, np.zeros(n))
            assert np.array_equal(U, g[f"ch{i}_U"]) and np.array_equal(inv, g[f"ch{i}_inv"])
    codes = set()
    for i in range(int(g["qp_count"][0])):
        n = int(g[f"qp{i}_n"][0])
        x = g[f"qp{i}_x0"].copy(); cl = np.zeros(n, np.int32); nf = np.zeros(1, np.int32); inv = np.zeros(sym(n))
        ret = lib.boxQP(g[f"qp{i}_H"].copy(), g[f"qp{i}_g"], g[f"qp{i}_lo"], g[f"qp{i}_hi"], x, np.zeros(sym(n)), np.zeros(sym(n)),
                        np.zeros(n), np.zeros(n), np.zeros(n), cl, nf, inv, n)
        codes.add(ret)
        assert ret == int(g[f"qp{i}_ret"][0]), f"boxQP case {i}"
        assert np.array_equal(x, g[f"qp{i}_x"]) and np.array_equal(cl, g[f"qp{i}_clamped"]) and nf[0] == g[f"qp{i}_nfree"][0]
        if ret >= 1 and ret != 6:
            assert np.array_equal(inv, g[f"qp{i}_inv"])
    assert {-1, 5, 6} <= codes, f"fixture should cover not-PD, converged and all-clamped exits, got {codes}"


def test_setoptparam_messages():
    """Option validation of the port: every message of the reference (iLQG.c:80-89) on its trigger."""
    s = oracle_lib.OracleLib("port", "car", 0).solver(4)
    cases = [("alpha", [1.5, 0.5], "all alpha must be in the range [1.0..0.0)"), ("alpha", [0.5, 0.5], "all alpha must be monotonically decreasing"),
             ("tolFun", [1.0, 2.0], "parameter must be scalar"), ("tolFun", 0.0, "parameter must be positive"), ("max_iter", -1, "parameter must be positive"),
             ("lambdaFactor", 0.5, "parameter must be > 1"), ("regType", 3, "parameter must be in range [1..2]"), ("zMin", 1.0, "parameter must be in range [0..1)"),
             ("debug_level", 7, "parameter must be in range [0..6]"), ("w_pen_init", 40, "no such parameter"), ("w_pen_fact2", 2.0, None), ("max_iter", 0, None)]
    refs = oracle_lib.OracleLib("reference", "car", 0).solver(4) if oracle_lib.available("reference", "car", 0) else None
    for name, v, msg in cases:
        assert s.set_opt_raw(name, v) == msg, name
        if refs is not None:
            assert refs.set_opt_raw(name, v) == msg, name


def test_inverse_columns_by_lane_equal_sequential_inverse(tmp_path):
    """The warp-cooperative backward pass computes the columns of the box QP's explicit inverse in different lanes
    (box_qp<M, CL> in csrc/ilqg_kernels.cuh); the restructured loops must give the reference's bits (cholesky.c:51-74)."""
    import subprocess

    exe = tmp_path / "inverse_columns_check"
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "inverse_columns_check.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", str(exe), src], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
