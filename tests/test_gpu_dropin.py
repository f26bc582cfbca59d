"""The single-problem drop-in (csrc/ilqg_dropin.c): the reference's own entry points -- standard_parameters,
setOptParam, init_opt, forward_pass, makeCandidateNominal, iLQG on a caller-allocated tOptSet -- executed on the GPU.
The SAME harness (oracle/harness.c, a stand-in for iLQG_mex.c) is linked once against the unmodified reference and
once against libilqg_dropin_*.so; everything the reference leaves in the tOptSet must be bit-identical."""
import numpy as np
import pytest

import oracle_lib
import parity_util as PU
from ilqg_b200 import workloads as W

pytestmark = pytest.mark.gpu
FIELDS = ("x", "u", "l", "L", "mult_f", "mult_t")
SCALARS = ("cost", "new_cost", "dcost", "expected", "lambda", "g_norm", "iterations", "dV0", "dV1", "w_pen_l", "w_pen_f", "n_linesearch")


def run(kind, problem, ddp, T, params, x0, u0, opts):
    s = oracle_lib.OracleLib(kind, problem, ddp).solver(T)
    s.set_opts(opts)
    s.set_params(params)
    rec = {"init": s.init(x0, u0)}
    rec["cost0"] = s.scalar("cost")
    rec["x_init"] = s.get("x")
    rec["result"] = s.solve()
    for k in SCALARS:
        rec[k] = s.scalar(k)
    for k in FIELDS:
        rec[k] = s.get(k)
    rec["log_linesearch"] = s.get("log_linesearch")
    s.close()
    return rec


def compare(problem, ddp, T, params, x0, u0, opts):
    kind = PU.oracle_kinds(problem, ddp)[0]
    assert oracle_lib.available("b200", problem, ddp), "build the drop-in harness: make -C oracle b200"
    a, b = run(kind, problem, ddp, T, params, x0, u0, opts), run("b200", problem, ddp, T, params, x0, u0, opts)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), f"{problem} ddp{ddp}: {k} differs between {kind} and the GPU drop-in"


@pytest.mark.parametrize("ddp", [0, 1])
def test_car_through_reference_api(ddp):
    x0, u0 = W.car_single()
    compare("car", ddp, 500, W.CAR_PARAMS, x0, u0, {"max_iter": 200})


@pytest.mark.parametrize("n", [2, 5, 500])
def test_brachi_through_reference_api(n):
    params, x0, u0, opts = W.brachi(n)
    compare("brachi", 0, n, params, x0, u0, opts)


def test_brachi_hli_through_reference_api():
    params, x0, u0, opts = W.brachi_hli(500)
    compare("brachi_hli", 0, 500, params, x0, u0, opts)


@pytest.mark.parametrize("b,ddp", [(0, 0), (1, 1), (3, 1)])
def test_pend_through_reference_api(b, ddp):
    """running equality + terminal inequality multipliers through the tOptSet marshalling (b=1/ddp=1 is a failed solve)"""
    x0, u0 = W.pend_batch(4)
    compare("pend", ddp, W.PEND_T, W.PEND_PARAMS, x0[b], u0[b], W.PEND_OPTS)


def test_carhx_through_reference_api():
    x0, u0 = W.car_single()
    compare("carhx", 0, 500, W.CARHX_PARAMS, x0, u0, {"max_iter": 40})


def test_option_errors_through_reference_api():
    s = oracle_lib.OracleLib("b200", "car", 0).solver(4)
    assert s.set_opt_raw("zMin", 1.0) == "parameter must be in range [0..1)"
    assert s.set_opt_raw("w_pen_init", 40.0) == "no such parameter"
    assert s.set_opt_raw("alpha", [1.0, 0.5, 0.25]) is None
