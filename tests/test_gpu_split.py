"""k_backpass_split (four lanes per problem, the small-batch backward pass) against k_backpass (one lane per problem) and
against the oracle.  Small batches pick the split kernel by default, so the other GPU parity tests already run through it;
here both kernels are forced in turn on the same inputs and EVERYTHING the backward pass writes must be bit-identical:
gains L, feed-forward l, dV, g_norm, the per-step box-QP active sets and return codes, and with them the whole solve.
Reference: back_pass.c:38-257, boxQP.c:39-238, iLQG.c:261-303."""
import numpy as np
import pytest

import parity_util as PU
from ilqg_b200 import workloads as W

pytestmark = pytest.mark.gpu

KEYS = ("result", "iterations", "n_ls", "n_bp", "cost", "lambda", "g_norm", "w_pen_l", "w_pen_f", "dV0", "dV1", "tr_alpha",
        "tr_lambda", "tr_newcost", "tr_clamp", "x", "u", "l", "L")


def both(problem, T, params, x0, u0, opts, ddp=0):
    lane = PU.gpu_records(problem, ddp, T, params, x0, u0, opts, tuning={"bp_split": 0})
    split = PU.gpu_records(problem, ddp, T, params, x0, u0, opts, tuning={"bp_split": 4})
    for b, (a, s) in enumerate(zip(lane, split)):
        for k in KEYS:
            assert np.array_equal(np.asarray(a[k]), np.asarray(s[k]), equal_nan=True), f"{problem} ddp{ddp} b{b}: {k} differs between the kernels"
    return split


@pytest.mark.parametrize("opts", [{"max_iter": 60}, {"max_iter": 25, "regType": 2}, {"max_iter": 25, "lambdaInit": 1e-8, "lambdaMin": 1e-12},
                                   {"max_iter": 20, "lambdaInit": 50.0, "lambdaMax": 400.0}])
def test_car_split_equals_lane_and_oracle(opts):
    """37 problems = 4 full warps of 8 problems and a ragged fifth; long enough that the box QP clamps, backtracks and the
    regularisation loop retries failed passes."""
    B, T = 37, 150
    x0, u0 = W.car_batch(B, T=T, seed=11)
    split = both("car", T, W.CAR_PARAMS, x0, u0, opts)
    kind = PU.oracle_kinds("car", 0)[0]
    for b in range(0, B, 3):
        ora = PU.oracle_record(kind, "car", 0, T, W.CAR_PARAMS, x0[b], u0[b], opts)
        PU.assert_same(split[b], ora, f"car split b{b} vs {kind}", keys=PU.assert_same.__defaults__[0] + ("l", "L", "dV0", "dV1"))


@pytest.mark.parametrize("opts", [{"max_iter": 45}, {"max_iter": 20, "regType": 2}, {"max_iter": 20, "lambdaInit": 1e-8, "lambdaMin": 1e-12}])
def test_car_full_ddp_split_equals_lane_and_oracle(opts):
    """FULL_DDP = 1: the second-order tensor terms (back_pass.c:95-131) through the split kernel; includes failed back passes
    (the car with FULL_DDP needs the regularisation retries)."""
    B, T = 21, 150
    x0, u0 = W.car_batch(B, T=T, seed=13)
    split = both("car", T, W.CAR_PARAMS, x0, u0, opts, ddp=1)
    kind = PU.oracle_kinds("car", 1)[0]
    for b in range(0, B, 4):
        ora = PU.oracle_record(kind, "car", 1, T, W.CAR_PARAMS, x0[b], u0[b], opts)
        PU.assert_same(split[b], ora, f"car ddp1 split b{b} vs {kind}", keys=PU.assert_same.__defaults__[0] + ("l", "L", "dV0", "dV1"))


def test_car_full_horizon_split():
    """BASELINE config 3 subset through the split kernel, T = 500, max_iter = 50, against the oracle."""
    B = 24
    x0, u0 = W.car_batch(B)
    opts = {"max_iter": 50}
    split = both("car", 500, W.CAR_PARAMS, x0, u0, opts)
    kind = PU.oracle_kinds("car", 0)[0]
    for b in range(0, B, 5):
        ora = PU.oracle_record(kind, "car", 0, 500, W.CAR_PARAMS, x0[b], u0[b], opts)
        PU.assert_same(split[b], ora, f"car split b{b} vs {kind}")


@pytest.mark.parametrize("ddp", [0, 1])
@pytest.mark.parametrize("n", [2, 3, 5, 500])
def test_brachistochrone_split(n, ddp):
    """State dimension below the lane count: lanes without a column of their own idle."""
    params, x0, u0, opts = W.brachi(n)
    split = both("brachi", n, params, x0[None], u0[None], opts, ddp=ddp)
    for kind in PU.oracle_kinds("brachi", ddp):
        PU.assert_same(split[0], PU.oracle_record(kind, "brachi", ddp, n, params, x0, u0, opts), f"brachi n={n} ddp{ddp} split vs {kind}")


def test_problems_with_multipliers_split():
    """Augmented-Lagrangian problems (multipliers, penalty weights): pendulum (hle + hfi) and Brachistochrone with a running
    inequality; the failed solve of the pendulum set is included."""
    x0, u0 = W.pend_batch(9)
    for ddp in (0, 1):
        both("pend", W.PEND_T, W.PEND_PARAMS, x0, u0, W.PEND_OPTS, ddp=ddp)
    params, x0, u0, opts = W.brachi_hli(120)
    for ddp in (0, 1):
        both("brachi_hli", 120, params, x0[None], u0[None], opts, ddp=ddp)


@pytest.mark.parametrize("tuning", [{"bp_split": 4, "bp_ppw": 8}, {"bp_split": 4, "bp_ppw": 2}, {"bp_split": 4, "bp_ppw": 1},
                                     {"bp_split": 0, "bp_ppw": 16}, {"bp_split": 0, "bp_ppw": 3}, {"bp_split": 0, "bp_latency": 0, "bp_ppw": 8}])
def test_problems_per_warp_do_not_change_results(tuning):
    """Sparse warps (fewer problems per warp, both kernels, both register builds of k_backpass): the same records as the default."""
    B, T = 21, 90
    x0, u0 = W.car_batch(B, T=T, seed=19)
    opts = {"max_iter": 30}
    ref = PU.gpu_records("car", 0, T, W.CAR_PARAMS, x0, u0, opts, tuning={"bp_split": 0, "bp_ppw": 32})
    got = PU.gpu_records("car", 0, T, W.CAR_PARAMS, x0, u0, opts, tuning=tuning)
    for b in range(B):
        for k in KEYS:
            assert np.array_equal(np.asarray(ref[b][k]), np.asarray(got[b][k]), equal_nan=True), f"{tuning} b{b}: {k}"


def test_split_is_the_default_for_small_batches_only():
    import ilqg_b200

    x0, u0 = W.car_batch(16, T=40, seed=5)
    for tuning, want in (({}, "k_backpass_split"), ({"bp_split": 0}, "k_backpass")):
        s = ilqg_b200.BatchSolver("car", 0, 16, 40, )
        s.set_options({"max_iter": 3}); s.set_params(W.CAR_PARAMS)
        for k, v in tuning.items():
            s.set_tuning(k, v)
        s.solve(x0, u0)
        assert s.get_int("bp_split")[0] == (4 if want == "k_backpass_split" else 0), want
        s.close()
