"""CUDA path vs the committed golden fixtures (recorded from the unmodified reference core): no oracle library
needed at run time.  Bit-exact."""
import os

import numpy as np
import pytest

import parity_util as PU
from ilqg_b200 import workloads as W

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = ("result", "iterations", "n_ls", "n_bp", "cost", "lambda", "g_norm", "w_pen_l", "w_pen_f", "dV0", "dV1",
        "tr_alpha", "tr_lambda", "tr_newcost", "x", "u", "l", "L")


def test_car_golden_batch():
    """All car fixtures with T=100 as ONE batch: ragged convergence inside a batch must not disturb any problem."""
    gs = [np.load(os.path.join(GOLD, f"car_T100_b{b}.npz")) for b in range(8)]
    x0 = np.stack([g["x0"] for g in gs])
    u0 = np.stack([g["u0"] for g in gs])
    recs = PU.gpu_records("car", 0, 100, W.CAR_PARAMS, x0, u0, {"max_iter": 30})
    for b, g in enumerate(gs):
        PU.assert_same(recs[b], {k: g[k] for k in KEYS}, f"car_T100_b{b}", keys=KEYS)


@pytest.mark.parametrize("ddp", [0, 1])
def test_car_single_golden(ddp):
    g = np.load(os.path.join(GOLD, f"car_single_ddp{ddp}.npz"))
    rec = PU.gpu_records("car", ddp, 500, W.CAR_PARAMS, g["x0"], g["u0"], {"max_iter": 200})[0]
    PU.assert_same(rec, {k: g[k] for k in KEYS}, f"car_single_ddp{ddp}", keys=KEYS)
    if (g["qp_ret_last"] >= 1).all():
        code = rec["tr_clamp"]
        assert np.array_equal(np.stack([code & 3, (code >> 2) & 3], axis=1)[::-1], g["qp_clamped_last"])
        assert np.array_equal(((code >> 16) & 0xff)[::-1], g["qp_ret_last"])


@pytest.mark.parametrize("ddp", [0, 1])
def test_car_state_dependent_limits_golden(ddp):
    """State-dependent input constraints (hx != 0): the active constraint's gradient enters the gains of clamped inputs."""
    g = np.load(os.path.join(GOLD, f"carhx_ddp{ddp}.npz"))
    rec = PU.gpu_records("carhx", ddp, 500, W.CARHX_PARAMS, g["x0"], g["u0"], {"max_iter": 60})[0]
    PU.assert_same(rec, {k: g[k] for k in KEYS}, f"carhx_ddp{ddp}", keys=KEYS)
    if (g["qp_ret_last"] >= 1).all():
        code = rec["tr_clamp"]
        assert np.array_equal(np.stack([code & 3, (code >> 2) & 3], axis=1)[::-1], g["qp_clamped_last"])


@pytest.mark.parametrize("ddp", [0, 1])
def test_brachi_running_inequality_golden(ddp):
    """hli running inequality (Ruxton multiplier update), terminal equality and a [k]-indexed parameter."""
    import ilqg_b200
    g = np.load(os.path.join(GOLD, f"brachi_hli_ddp{ddp}.npz"))
    params, _, _, opts = W.brachi_hli(500)
    rec = PU.gpu_records("brachi_hli", ddp, 500, params, g["x0"][None], g["u0"][None], opts)[0]
    PU.assert_same(rec, {k: g[k] for k in KEYS}, f"brachi_hli_ddp{ddp}", keys=KEYS)
    s = ilqg_b200.BatchSolver("brachi_hli", ddp, 1, 500)
    s.set_options(opts); s.set_params(params); s.solve(g["x0"][None], g["u0"][None])
    assert s.get("mu_f").ravel()[0] == g["mult_f"][0]
    assert np.array_equal(s.get("mu_r").reshape(500), g["mult_t"][:, 0])      # running multipliers mu_li[k]
    s.close()


@pytest.mark.parametrize("ddp", [0, 1])
def test_pendulum_running_equality_and_terminal_inequality_golden(ddp):
    """hle running equality + hfi terminal inequality + torque limits, four problems as one batch (one of them fails for
    FULL_DDP=1): solver record, box-QP active sets, running multipliers mu_le[k] / last_hle[k] and final mu_fi / last_hfi."""
    import ilqg_b200
    gs = [np.load(os.path.join(GOLD, f"pend_b{b}_ddp{ddp}.npz")) for b in range(4)]
    x0, u0 = np.stack([g["x0"] for g in gs]), np.stack([g["u0"] for g in gs])
    recs = PU.gpu_records("pend", ddp, W.PEND_T, W.PEND_PARAMS, x0, u0, W.PEND_OPTS)
    for b, g in enumerate(gs):
        PU.assert_same(recs[b], {k: g[k] for k in KEYS}, f"pend_b{b}_ddp{ddp}", keys=KEYS)
        if (g["qp_ret_last"] >= 1).all():
            code = recs[b]["tr_clamp"]
            assert np.array_equal(np.stack([code & 3, (code >> 2) & 3], axis=1)[::-1], g["qp_clamped_last"])
    s = ilqg_b200.BatchSolver("pend", ddp, 4, W.PEND_T)
    s.set_options(W.PEND_OPTS); s.set_params(W.PEND_PARAMS); s.solve(x0, u0)
    mu_f, last_f, mu_r, last_r = s.get("mu_f"), s.get("last_f"), s.get("mu_r"), s.get("last_r")
    for b, g in enumerate(gs):
        assert mu_f[b].ravel()[0] == g["mult_f"][0] and last_f[b].ravel()[0] == g["mult_f"][1], b      # mu_fi, last_hfi
        assert np.array_equal(mu_r[b].reshape(W.PEND_T), g["mult_t"][:, 0]), b                          # mu_le[k]
        assert np.array_equal(last_r[b].reshape(W.PEND_T), g["mult_t"][:, 1]), b                        # last_hle[k]
    s.close()


@pytest.mark.parametrize("ddp", [0, 1])
def test_quad_golden(ddp):
    gs = [np.load(os.path.join(GOLD, f"quad_T300_b{b}_ddp{ddp}.npz")) for b in range(2)]
    recs = PU.gpu_records("quad", ddp, 300, W.QUAD_PARAMS, np.stack([g["x0"] for g in gs]), np.stack([g["u0"] for g in gs]), {"max_iter": 25})
    for b, g in enumerate(gs):
        PU.assert_same(recs[b], {k: g[k] for k in KEYS}, f"quad_T300_b{b}_ddp{ddp}", keys=KEYS)


@pytest.mark.parametrize("n", [2, 3, 5, 500])
@pytest.mark.parametrize("ddp", [0, 1])
def test_brachi_golden(n, ddp):
    g = np.load(os.path.join(GOLD, f"brachi_n{n}_ddp{ddp}.npz"))
    params, _, _, opts = W.brachi(n)
    rec = PU.gpu_records("brachi", ddp, n, params, g["x0"][None], g["u0"][None], opts)[0]
    PU.assert_same(rec, {k: g[k] for k in KEYS}, f"brachi n={n} ddp{ddp}", keys=KEYS)
    import ilqg_b200  # final multiplier mu_fe and its trace end state
    s = ilqg_b200.BatchSolver("brachi", ddp, 1, n)
    s.set_options(opts); s.set_params(params); s.solve(g["x0"][None], g["u0"][None])
    assert s.get("mu_f").ravel()[0] == g["mult_f"][0]
    s.close()
