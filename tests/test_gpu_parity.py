"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on identical inputs.
Bar: BIT-EXACT -- identical iteration count, accepted-alpha / lambda sequences, box-QP active sets, and
bit-identical final cost and trajectories (stricter than the 1e-9 relative of BASELINE.json)."""
import numpy as np
import pytest

import parity_util as PU
from ilqg_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ddp", [0, 1])
def test_car_single_matches_oracle(ddp):
    """BASELINE config 1: car, T=500, max_iter=200, single instance (here problem 0 of a batch of 3)."""
    x0, u0 = W.car_single()
    xb, ub = W.car_batch(2, seed=7)
    X0 = np.vstack([x0[None], xb])
    U0 = np.concatenate([u0[None], ub])
    opts = {"max_iter": 200}
    recs = PU.gpu_records("car", ddp, 500, W.CAR_PARAMS, X0, U0, opts)
    kinds = PU.oracle_kinds("car", ddp)
    assert kinds, "no oracle library built"
    for kind in kinds:
        for b in range(3):
            ora = PU.oracle_record(kind, "car", ddp, 500, W.CAR_PARAMS, X0[b], U0[b], opts, qp_cap=400000)
            PU.assert_same(recs[b], ora, f"car ddp{ddp} b{b} vs {kind}")
            # active sets of the last back pass: the last T box-QP calls of the oracle, k = T-1 .. 0
            if ora["result"] == 1 or True:
                ret, nfree, cl = ora["qp"]
                if len(ret) >= 500 and (ret[-500:] >= 1).all():
                    code = recs[b]["tr_clamp"]
                    gpu_cl = np.stack([(code >> 0) & 3, (code >> 2) & 3], axis=1)[::-1]
                    gpu_ret = ((code >> 16) & 0xff)[::-1]
                    assert np.array_equal(gpu_cl, cl[-500:]), f"active sets differ ({kind}, b{b})"
                    assert np.array_equal(gpu_ret, ret[-500:]), f"QP return codes differ ({kind}, b{b})"


@pytest.mark.parametrize("n", [2, 3, 5, 500])
@pytest.mark.parametrize("ddp", [0, 1])
def test_brachistochrone_matches_oracle(n, ddp):
    """BASELINE config 2: Brachistochrone, all four horizons of testBrachi.m:18."""
    params, x0, u0, opts = W.brachi(n)
    recs = PU.gpu_records("brachi", ddp, n, params, x0[None], u0[None], opts)
    for kind in PU.oracle_kinds("brachi", ddp):
        ora = PU.oracle_record(kind, "brachi", ddp, n, params, x0, u0, opts)
        PU.assert_same(recs[0], ora, f"brachi n={n} ddp{ddp} vs {kind}")


def test_car_batch_matches_oracle():
    """BASELINE config 3/4 parity subset: first 64 problems of the synthetic batch, max_iter=50."""
    B = 64
    x0, u0 = W.car_batch(B)
    opts = {"max_iter": 50}
    recs = PU.gpu_records("car", 0, 500, W.CAR_PARAMS, x0, u0, opts)
    kind = PU.oracle_kinds("car", 0)[0]
    for b in range(B):
        ora = PU.oracle_record(kind, "car", 0, 500, W.CAR_PARAMS, x0[b], u0[b], opts)
        PU.assert_same(recs[b], ora, f"car batch b{b} vs {kind}")


@pytest.mark.parametrize("ddp", [1, 0])
def test_quadrotor_matches_oracle(ddp):
    """BASELINE config 5 parity subset: synthetic quadrotor n=12, m=4, T=1000 (warp-cooperative backward pass; FULL_DDP=1
    exercises the regularisation retry loop: ~44 back passes for 24 passes)."""
    B, T = 4, 1000
    x0, u0 = W.quad_batch(B, T=T)
    opts = {"max_iter": 30 if ddp else 12}
    recs = PU.gpu_records("quad", ddp, T, W.QUAD_PARAMS, x0, u0, opts)
    kind = PU.oracle_kinds("quad", ddp)[0]
    for b in range(B):
        ora = PU.oracle_record(kind, "quad", ddp, T, W.QUAD_PARAMS, x0[b], u0[b], opts)
        PU.assert_same(recs[b], ora, f"quad ddp{ddp} b{b} vs {kind}", keys=("result", "iterations", "n_ls", "n_bp", "cost", "lambda", "g_norm",
                                                                            "tr_alpha", "tr_lambda", "tr_newcost", "x", "u", "l", "L", "dV0", "dV1"))
