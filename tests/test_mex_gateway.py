"""The MATLAB / Octave gateway: ddp-generator_b200/mex/iLQG_mex_b200.c against the reference's own
gateway iLQG_mex.c, both driven in-process through the fake mex API of oracle/mex_stub/fake_mex.c.

CPU part: the reference gateway over the fake API reproduces the golden fixtures (pins the test rig itself), and every
argument error of the reference gateway comes back from ours with the same identifier and text -- all argument checking
happens before the GPU is touched.  GPU part: same call, same outputs, bit for bit; plus the batched extension."""
import os

import numpy as np
import pytest

import fake_mex as FM
from ilqg_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_rig = pytest.mark.skipif(not os.path.exists(FM.FAKEMEX), reason="build the fake mex runtime: make -C oracle port")


def have(kind, problem, ddp=0):
    return os.path.exists(FM.gateway_path(kind, problem, ddp))


def gateway(kind, problem, ddp=0):
    if not have(kind, problem, ddp):
        if kind == "reference":
            pytest.skip("reference gateway not built here (oracle/_ref needs /root/reference)")
        pytest.fail("build the gateway over the GPU library: make -C oracle b200")
    return FM.Gateway(FM.gateway_path(kind, problem, ddp))


# ---------------------------------------------------------------------------------------------------- CPU
@needs_rig
@pytest.mark.parametrize("ddp", [0, 1])
def test_reference_gateway_reproduces_golden(ddp):
    g = gateway("reference", "car", ddp)
    x0, u0 = W.car_single()
    ok, x, u, cost = g(x0, u0.T, W.CAR_PARAMS, {"max_iter": 200.0})
    gold = np.load(os.path.join(GOLD, f"car_single_ddp{ddp}.npz"))
    assert ok.shape == (1, 1) and x.shape == (4, 501) and u.shape == (2, 500) and cost.shape == (1, 1)
    assert ok[0, 0] == gold["result"] and cost[0, 0] == gold["cost"]
    assert np.array_equal(x.T, gold["x"]) and np.array_equal(u.T, gold["u"])


def _car_args():
    x0, u0 = W.car_single(T=20)
    return [x0, u0.T, dict(W.CAR_PARAMS), {"max_iter": 5.0}]


def _mutations():
    def m(name, fn, nlhs=4):
        return pytest.param(fn, nlhs, id=name)

    def drop_arg(a):
        return a[:3]

    def states(a):
        a[0] = np.zeros(5)
        return a

    def inputs(a):
        a[1] = np.zeros((3, 20))
        return a

    def params_not_struct(a):
        a[2] = np.zeros(3)
        return a

    def opt_not_struct(a):
        a[3] = np.zeros(3)
        return a

    def opt(name, value):
        def f(a):
            a[3] = {name: value}
            return a
        return f

    def missing_param(a):
        del a[2]["limW"]
        return a

    def param_len(a):
        a[2]["pf"] = [1.0, 2.0]
        return a

    def param_matrix(a):
        a[2]["pf"] = np.ones((2, 2))
        return a

    def param_sparse(a):
        a[2]["d"] = FM.Sparse([2.0])
        return a

    return [
        m("three_inputs", drop_arg), m("three_outputs", lambda a: a, nlhs=3), m("wrong_states", states), m("wrong_inputs", inputs),
        m("params_not_struct", params_not_struct), m("opt_not_struct", opt_not_struct),
        m("alpha_not_decreasing", opt("alpha", [1.0, 1.0])), m("alpha_out_of_range", opt("alpha", [1.5, 0.5])),
        m("unknown_option", opt("w_pen_init", 1.0)), m("regType_3", opt("regType", 3.0)), m("zMin_1", opt("zMin", 1.0)),
        m("lambdaFactor_small", opt("lambdaFactor", 0.5)), m("option_not_scalar", opt("tolFun", [1.0, 2.0])),
        m("missing_param", missing_param), m("param_wrong_length", param_len), m("param_matrix", param_matrix),
        m("param_sparse", param_sparse),
    ]


@needs_rig
@pytest.mark.parametrize("mutate,nlhs", _mutations())
def test_argument_errors_match_reference_gateway(mutate, nlhs):
    ours, ref = gateway("b200", "car"), gateway("reference", "car")
    with pytest.raises(FM.MexError) as e_ref:
        ref(*mutate(_car_args()), nlhs=nlhs)
    with pytest.raises(FM.MexError) as e_ours:
        ours(*mutate(_car_args()), nlhs=nlhs)
    assert (e_ours.value.ident, e_ours.value.msg) == (e_ref.value.ident, e_ref.value.msg)


@needs_rig
def test_gateway_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(FM.MexError) as e:
        gateway("b200", "car")(*_car_args())
    assert e.value.ident == "iLQG:gpu" and e.value.msg


# ---------------------------------------------------------------------------------------------------- GPU
def _same_as_reference(problem, ddp, x0, u0, params, opts):
    ours = gateway("b200", problem, ddp)(x0, u0.T, params, opts)
    if have("reference", problem, ddp):
        ref = gateway("reference", problem, ddp)(x0, u0.T, params, opts)
        for a, b, what in zip(ours, ref, ("success", "x", "u", "cost")):
            assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), f"{problem} ddp{ddp}: {what} differs from iLQG_mex.c"
    return ours


@pytest.mark.gpu
@pytest.mark.parametrize("ddp", [0, 1])
def test_gpu_gateway_car_single(ddp):
    x0, u0 = W.car_single()
    ok, x, u, cost = _same_as_reference("car", ddp, x0, u0, W.CAR_PARAMS, {"max_iter": 200.0})
    gold = np.load(os.path.join(GOLD, f"car_single_ddp{ddp}.npz"))
    assert ok[0, 0] == gold["result"] and cost[0, 0] == gold["cost"]
    assert np.array_equal(x.T, gold["x"]) and np.array_equal(u.T, gold["u"])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 5, 500])
def test_gpu_gateway_brachi(n):
    params, x0, u0, opts = W.brachi(n)
    ok, x, u, cost = _same_as_reference("brachi", 0, x0, u0, params, opts)
    gold = np.load(os.path.join(GOLD, f"brachi_n{n}_ddp0.npz"))
    assert cost[0, 0] == gold["cost"] and np.array_equal(x.T, gold["x"]) and np.array_equal(u.T, gold["u"])


@pytest.mark.gpu
def test_gpu_gateway_k_indexed_parameter():
    params, x0, u0, opts = W.brachi_hli(500)          # ymin has one value per timestep (length N)
    ok, x, u, cost = _same_as_reference("brachi_hli", 0, x0, u0, params, opts)
    gold = np.load(os.path.join(GOLD, "brachi_hli_ddp0.npz"))
    assert cost[0, 0] == gold["cost"] and np.array_equal(x.T, gold["x"])


@pytest.mark.gpu
def test_gpu_gateway_batched():
    """u_nom m x (N-1) x B: one call solves what B reference calls solve."""
    B, T = 8, 100
    xb, ub = W.car_batch(B, T=T, seed=5)
    ok, x, u, cost = gateway("b200", "car")(xb.T, np.transpose(ub, (2, 1, 0)), W.CAR_PARAMS, {"max_iter": 30.0})
    assert ok.shape == (1, B) and x.shape == (4, T + 1, B) and u.shape == (2, T, B) and cost.shape == (1, B)
    for b in range(B):
        gold = np.load(os.path.join(GOLD, f"car_T100_b{b}.npz"))
        assert ok[0, b] == gold["result"] and cost[0, b] == gold["cost"], b
        assert np.array_equal(x[:, :, b].T, gold["x"]) and np.array_equal(u[:, :, b].T, gold["u"]), b


@pytest.mark.gpu
def test_gpu_gateway_parameter_column_per_problem():
    """A k x B parameter matrix gives every problem its own parameter vector (B independent reference calls)."""
    B, T = 4, 60
    xb, ub = W.car_batch(B, T=T, seed=9)
    cu = np.array([[1e-2, 2e-2, 5e-3, 1e-2], [1e-4, 1e-4, 3e-4, 5e-5]])
    lim = np.array([[-0.5, -0.3, -0.5, -0.2], [0.5, 0.3, 0.4, 0.2]])
    params = dict(W.CAR_PARAMS, cu=cu, limW=lim)
    opts = {"max_iter": 25.0}
    ok, x, u, cost = gateway("b200", "car")(xb.T, np.transpose(ub, (2, 1, 0)), params, opts)
    if not have("reference", "car"):
        pytest.skip("per-problem comparison needs the reference gateway (oracle/_ref)")
    ref = gateway("reference", "car")
    for b in range(B):
        rok, rx, ru, rcost = ref(xb[b], ub[b].T, dict(W.CAR_PARAMS, cu=cu[:, b], limW=lim[:, b]), opts)
        assert ok[0, b] == rok[0, 0] and cost[0, b] == rcost[0, 0], b
        assert np.array_equal(x[:, :, b], rx) and np.array_equal(u[:, :, b], ru), b


@pytest.mark.gpu
def test_gpu_gateway_nonfinite_start():
    """A non-finite initial rollout: success 0, x and u left zero (iLQG_mex.c:116-118)."""
    x0, u0 = W.car_single(T=50)
    u0[7, 1] = np.nan
    ok, x, u, cost = _same_as_reference("car", 0, x0, u0, W.CAR_PARAMS, {"max_iter": 5.0})
    assert ok[0, 0] == 0 and not x.any() and not u.any()
