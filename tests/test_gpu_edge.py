"""Edge cases and size-independent properties of the CUDA path (through the C ABI)."""
import numpy as np
import pytest

import ilqg_b200
import oracle_lib
import parity_util as PU
from ilqg_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _oracle(problem, ddp):
    return oracle_lib.OracleLib(PU.oracle_kinds(problem, ddp)[0], problem, ddp)


@pytest.mark.parametrize("B", [1, 31, 33, 65])
def test_ragged_batch_sizes(B):
    x0, u0 = W.car_batch(B, T=60, seed=21)
    opts = {"max_iter": 8}
    s = ilqg_b200.BatchSolver("car", 0, B, 60)
    s.set_options(opts); s.set_params(W.CAR_PARAMS)
    out = s.solve(x0, u0)
    ora = _oracle("car", 0).solve_batch(x0, u0, W.CAR_PARAMS, {"max_iter": 8.0}, 2, want_traj=True)
    for k in ("cost", "iterations", "n_linesearch", "x", "u"):
        assert np.array_equal(out[k], ora[k]), k
    assert np.array_equal(out["success"], ora["result"])
    s.close()


@pytest.mark.parametrize("opts", [{"max_iter": 0}, {"max_iter": 1}, {"max_iter": 12, "zMin": 0.5}, {"max_iter": 12, "alpha": [1.0, 0.1]},
                                   {"max_iter": 12, "alpha": [0.5]}, {"max_iter": 15, "lambdaInit": 100.0, "lambdaMax": 1000.0},
                                   {"max_iter": 12, "lambdaFactor": 3.0, "lambdaMin": 1e-3, "dlambdaInit": 2.0},
                                   {"max_iter": 40, "tolFun": 1e-2}, {"max_iter": 40, "tolGrad": 1.0, "lambdaInit": 1e-6},
                                   {"max_iter": 12, "regType": 2}, {"max_iter": 10, "alpha": [1.0, 0.5, 0.25, 0.125, 0.0625, 0.03, 0.01, 0.005, 0.001, 0.0005]}])
def test_options_follow_reference(opts):
    B, T = 6, 80
    x0, u0 = W.car_batch(B, T=T, seed=31)
    recs = PU.gpu_records("car", 0, T, W.CAR_PARAMS, x0, u0, opts)
    kind = PU.oracle_kinds("car", 0)[0]
    for b in range(B):
        if opts["max_iter"] == 0:
            continue   # backPassDone is uninitialised in the reference for max_iter == 0: result undefined
        ora = PU.oracle_record(kind, "car", 0, T, W.CAR_PARAMS, x0[b], u0[b], opts)
        PU.assert_same(recs[b], ora, f"opts {opts} b{b}")
    if opts["max_iter"] == 0:
        assert all(r["iterations"] == 0 and r["n_ls"] == 0 for r in recs)


def test_failed_initial_rollout_is_reported():
    """A rollout whose dynamics turn non-finite makes the harness/mex report failure before iLQG starts
    (iLQG_mex.c:116-118): result -1 here, and the problem takes no passes."""
    B, T = 4, 30
    x0, u0 = W.car_batch(B, T=T, seed=41)
    x0[2, 3] = 1e4    # speed so large that sqrt(d^2 - (h v sin w)^2) is NaN
    s = ilqg_b200.BatchSolver("car", 0, B, T)
    s.set_options({"max_iter": 5}); s.set_params(W.CAR_PARAMS)
    out = s.solve(x0, u0)
    O = _oracle("car", 0)
    for b in range(B):
        h = O.solver(T); h.set_opts({"max_iter": 5}); h.set_params(W.CAR_PARAMS)
        ok = h.init(x0[b], u0[b])
        assert (out["success"][b] == -1) == (not ok)
        if ok:
            h.solve()
            assert out["cost"][b] == h.scalar("cost") and out["iterations"][b] == h.scalar("iterations")
    assert out["n_linesearch"][2] == 0
    s.close()


def test_all_controls_clamped_and_limits():
    """Controls far outside the box: the initial rollout clamps them and most QPs return 6."""
    B, T = 3, 50
    x0, u0 = W.car_batch(B, T=T, seed=51)
    u0 = u0 * 100.0
    recs = PU.gpu_records("car", 0, T, W.CAR_PARAMS, x0, u0, {"max_iter": 6})
    kind = PU.oracle_kinds("car", 0)[0]
    for b in range(B):
        ora = PU.oracle_record(kind, "car", 0, T, W.CAR_PARAMS, x0[b], u0[b], {"max_iter": 6})
        PU.assert_same(recs[b], ora, f"clamped b{b}")
        assert np.abs(recs[b]["u"][:, 0]).max() <= 0.5 and np.abs(recs[b]["u"][:, 1]).max() <= 2.0


def test_properties_at_config3_size():
    """BASELINE config 3 (B = 4096, T = 500): determinism, batch-composition invariance, monotone cost."""
    B, T = 4096, 500
    x0, u0 = W.car_batch(B, T=T)
    s = ilqg_b200.BatchSolver("car", 0, B, T, flags=ilqg_b200.TRACE)
    s.set_options({"max_iter": 12}); s.set_params(W.CAR_PARAMS)
    a = s.solve(x0, u0)
    tr_cost, tr_alpha = s.get("tr_newcost"), s.get_int("tr_alpha")
    cost0 = None
    b2 = s.solve(x0, u0)
    for k in ("cost", "iterations", "n_linesearch", "x", "u"):
        assert np.array_equal(a[k], b2[k]), f"rerun differs in {k}"
    # permuting the problems permutes the results
    perm = np.random.default_rng(0).permutation(B)
    c = s.solve(x0[perm], u0[perm])
    for k in ("cost", "iterations", "n_linesearch", "x", "u"):
        assert np.array_equal(c[k], a[k][perm]), f"permutation changes {k}"
    s.close()
    # a problem's result does not depend on which batch it is in
    sub = np.arange(100, 100 + 37)
    t = ilqg_b200.BatchSolver("car", 0, sub.size, T)
    t.set_options({"max_iter": 12}); t.set_params(W.CAR_PARAMS)
    d = t.solve(x0[sub], u0[sub])
    for k in ("cost", "iterations", "x", "u"):
        assert np.array_equal(d[k], a[k][sub]), k
    t.close()
    # every accepted step lowers the cost: accepted new_cost sequence is strictly decreasing per problem
    n_alpha = 8
    for b in range(0, B, 97):
        n = a["n_linesearch"][b]
        acc = tr_cost[b][:n][tr_alpha[b][:n] <= n_alpha]
        assert np.all(np.diff(acc) < 0)
    # and 64 of them against the oracle
    ora = _oracle("car", 0).solve_batch(x0[:64], u0[:64], W.CAR_PARAMS, {"max_iter": 12.0}, 4, want_traj=True)
    assert np.array_equal(ora["cost"], a["cost"][:64]) and np.array_equal(ora["x"], a["x"][:64])


@pytest.mark.parametrize("chunks", [2, 3, 5])
def test_chunked_streams_match_single_stream(chunks):
    """A handle split into several concurrent streams (ragged chunk sizes) returns exactly what one stream returns, for
    every read-back path (download, get, get_int), and matches the oracle."""
    B, T = 77, 60
    x0, u0 = W.car_batch(B, T=T, seed=61)
    opts = {"max_iter": 9}
    res = []
    for ch in (1, chunks):
        s = ilqg_b200.BatchSolver("car", 0, B, T, flags=ilqg_b200.TRACE, chunks=ch)
        assert 1 <= s.chunks() <= ch and (ch == 1 or s.chunks() > 1)    # chunk sizes are whole warps: fewer chunks may result
        s.set_options(opts); s.set_params(W.CAR_PARAMS)
        out = s.solve(x0, u0)
        out.update(l=s.get("l"), L=s.get("L"), lam=s.get("lambda"), tr=s.get_int("tr_alpha"), nbp=s.get_int("n_backpass"), v1=s.get("v1"))
        res.append(out)
        s.close()
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k
    ora = _oracle("car", 0).solve_batch(x0, u0, W.CAR_PARAMS, {"max_iter": 9.0}, 2, want_traj=True)
    assert np.array_equal(res[1]["cost"], ora["cost"]) and np.array_equal(res[1]["x"], ora["x"]) and np.array_equal(res[1]["u"], ora["u"])


def test_stepwise_iteration_and_caller_stream():
    """start + iterate(n) ... + finish on a caller-owned CUDA stream equals solve(); the handle's work is ordered in
    that stream (events recorded on it bracket the whole solve)."""
    torch = pytest.importorskip("torch")
    B, T = 40, 80
    x0, u0 = W.car_batch(B, T=T, seed=71)
    ref = ilqg_b200.BatchSolver("car", 0, B, T)
    ref.set_options({"max_iter": 10}); ref.set_params(W.CAR_PARAMS)
    want = ref.solve(x0, u0)
    ref.close()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        s = ilqg_b200.BatchSolver("car", 0, B, T, stream=st.cuda_stream, chunks=2)
        s.set_options({"max_iter": 10}); s.set_params(W.CAR_PARAMS)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.upload(x0, u0)
        e0.record(st)
        s.start()
        done = 0
        while done < 10:
            n = s.iterate(3)
            done += n
            if n == 0:
                break
        s.finish()
        e1.record(st)
        e1.synchronize()
        assert e0.elapsed_time(e1) > 0.0
        got = s.download()
        s.close()
    for k in ("cost", "iterations", "n_linesearch", "success", "x", "u"):
        assert np.array_equal(got[k], want[k]), k
    # the pipelined single-call path (upload + solve + download per chunk stream, descending stream priorities)
    p = ilqg_b200.BatchSolver("car", 0, B, T, chunks=2)
    p.set_options({"max_iter": 10}); p.set_params(W.CAR_PARAMS)
    piped = p.solve_host(x0, u0)
    p.close()
    for k in ("cost", "iterations", "n_linesearch", "success", "x", "u"):
        assert np.array_equal(piped[k], want[k]), k


def test_per_problem_parameter_sets():
    """Every problem of a batch with its own parameter struct (limits, cost weights, geometry), like B independent calls
    of the reference: bit-identical to solving each problem alone with its parameters."""
    B, T = 37, 70
    x0, u0 = W.car_batch(B, T=T, seed=81)
    rng = np.random.default_rng(5)
    limW = np.stack([-0.3 - 0.3 * rng.random(B), 0.3 + 0.3 * rng.random(B)], axis=1)
    cf = np.array(W.CAR_PARAMS["cf"])[None, :] * (0.5 + rng.random((B, 4)))
    d = 1.5 + rng.random((B, 1))
    opts = {"max_iter": 8}
    for chunks in (1, 2):
        s = ilqg_b200.BatchSolver("car", 0, B, T, chunks=chunks)
        s.set_options(opts); s.set_params(W.CAR_PARAMS)
        s.set_params_batch({"limW": limW, "cf": cf, "d": d})
        out = s.solve(x0, u0)
        s.close()
        O = _oracle("car", 0)
        for b in range(B):
            p = dict(W.CAR_PARAMS, limW=limW[b], cf=cf[b], d=d[b])
            h = O.solver(T); h.set_opts(opts); h.set_params(p)
            assert h.init(x0[b], u0[b]); h.solve()
            assert out["cost"][b] == h.scalar("cost") and out["iterations"][b] == h.scalar("iterations"), (chunks, b)
            assert np.array_equal(out["x"][b], h.get("x")) and np.array_equal(out["u"][b], h.get("u")), (chunks, b)


@pytest.mark.parametrize("tail_from", ["0", "1", "2", "8"])
def test_non_finite_rollouts_inside_the_line_search(tail_from, monkeypatch):
    """A regime where large-step rollouts leave the domain of the dynamics (sqrt of a negative number -> NaN guard ->
    forward_pass returns 0, line_search.c:55-59) while smaller steps are fine; all three line-search schedules
    (parallel tail after 1 or 2 rounds, purely sequential rounds) must replay the reference's decisions."""
    monkeypatch.setenv("ILQG_LS_TAIL_FROM", tail_from)
    B, T = 12, 80
    x0, u0 = W.car_batch(B, T=T, seed=91)
    x0[:, 3] = 1.5
    u0 = u0 * 5
    params = dict(W.CAR_PARAMS, d=[0.05], limA=[-20.0, 20.0])
    opts = {"max_iter": 15}
    recs = PU.gpu_records("car", 0, T, params, x0, u0, opts)
    kind = PU.oracle_kinds("car", 0)[0]
    deep = 0
    for b in range(B):
        ora = PU.oracle_record(kind, "car", 0, T, params, x0[b], u0[b], opts)
        PU.assert_same(recs[b], ora, f"nan-regime b{b} tail_from={tail_from}")
        deep += int((ora["tr_alpha"] >= 4).sum())
    assert deep > 20      # the case really exercises deep backtracking


def test_config4_full_size_sample_parity():
    """BASELINE config 4 at full size (262 144 car problems, T = 500; 4 concurrent chunks): problems next to every chunk
    boundary and at both ends match the oracle bit for bit; a rerun is bit-identical."""
    B, T, iters = 262144, 500, 4
    x0 = np.empty((B, 4)); u0 = np.empty((B, T, 2))
    for s0 in range(0, B, 32768):
        a, b = W.car_batch(32768, T=T, first=s0)
        x0[s0:s0 + 32768] = a; u0[s0:s0 + 32768] = b
    s = ilqg_b200.BatchSolver("car", 0, B, T)
    assert s.chunks() == 4
    s.set_options({"max_iter": iters}); s.set_params(W.CAR_PARAMS)
    out = s.solve(x0, u0, want_traj=False)
    cost_again = s.solve(x0, u0, want_traj=False)["cost"]
    assert np.array_equal(out["cost"], cost_again)
    idx = np.concatenate([np.arange(0, 8)] + [np.arange(c * 65536 - 3, c * 65536 + 3) for c in range(1, 4)] + [np.arange(B - 8, B)])
    xs = s.get("x")[idx]
    s.close()
    ora = _oracle("car", 0).solve_batch(x0[idx], u0[idx], W.CAR_PARAMS, {"max_iter": float(iters)}, 4, want_traj=True)
    assert np.array_equal(out["cost"][idx], ora["cost"]) and np.array_equal(out["iterations"][idx], ora["iterations"])
    assert np.array_equal(out["n_linesearch"][idx], ora["n_linesearch"]) and np.array_equal(xs, ora["x"])
    assert out["n_linesearch"].sum() == B * iters      # nobody converges within 4 passes on this workload


def test_config5_full_size_sample_parity():
    """BASELINE config 5 at full size (16 384 quadrotor problems, T = 1000, FULL_DDP = 1), a few passes."""
    B, T, iters = 16384, 1000, 5
    x0, u0 = W.quad_batch(B, T=T)
    s = ilqg_b200.BatchSolver("quad", 1, B, T)
    s.set_options({"max_iter": iters}); s.set_params(W.QUAD_PARAMS)
    out = s.solve(x0, u0, want_traj=False)
    idx = np.array([0, 1, 8191, 8192, B - 2, B - 1])
    us = s.get("u")[idx]
    s.close()
    ora = _oracle("quad", 1).solve_batch(x0[idx], u0[idx], W.QUAD_PARAMS, {"max_iter": float(iters)}, 3, want_traj=True)
    assert np.array_equal(out["cost"][idx], ora["cost"]) and np.array_equal(out["iterations"][idx], ora["iterations"])
    assert np.array_equal(us, ora["u"])
