"""Shared helpers for parity tests: run the oracle (reference or port) and the CUDA library on the same inputs and
collect comparable records."""
from __future__ import annotations

import numpy as np

import oracle_lib


def oracle_kinds(problem, ddp):
    """Checkers available on this machine: the compiled reference (oracle/_ref) and/or the C port."""
    return [k for k in ("reference", "port") if oracle_lib.available(k, problem, ddp)]


def oracle_record(kind, problem, ddp, T, params, x0, u0, opts, qp_cap=0):
    O = oracle_lib.OracleLib(kind, problem, ddp)
    s = O.solver(T)
    s.set_opts(opts)
    s.set_params(params)
    if qp_cap:
        s.qp_trace_enable(qp_cap)
    ok = s.init(x0, u0)
    rec = dict(init_ok=ok)
    if not ok:
        return rec
    rec["cost0"] = s.scalar("cost")
    rec["result"] = s.solve()
    rec["iterations"] = int(s.scalar("iterations"))
    for k in ("cost", "lambda", "g_norm", "w_pen_l", "w_pen_f", "dV0", "dV1"):
        rec[k] = s.scalar(k)
    for k in ("x", "u", "l", "L"):
        rec[k] = s.get(k)
    rec["mult_f"] = s.get("mult_f")
    rec["tr_lambda"] = s.trace("lambda")
    rec["tr_alpha"] = s.trace("alpha_idx").astype(int)
    rec["tr_newcost"] = s.trace("new_cost")
    rec["tr_iter"] = s.trace("iter").astype(int)
    rec["n_ls"] = rec["tr_lambda"].size
    rec["n_bp"] = s.bp_trace("lambda").size
    if qp_cap:
        rec["qp"] = s.qp_trace()
    s.close()
    return rec


def gpu_records(problem, ddp, T, params, x0, u0, opts, flags=1, tuning=None):
    """Solve a batch on the GPU; returns a list of per-problem records shaped like oracle_record()."""
    import ilqg_b200

    x0 = np.atleast_2d(x0)
    u0 = np.asarray(u0)
    if u0.ndim == 2:
        u0 = u0[None]
    B = x0.shape[0]
    s = ilqg_b200.BatchSolver(problem, ddp, B, T, flags=flags)
    s.set_options(opts)
    s.set_params(params)
    for name, value in (tuning or {}).items():
        s.set_tuning(name, value)
    out = s.solve(x0, u0)
    sc = {k: s.get(k) for k in ("lambda", "g_norm", "w_pen_l", "w_pen_f", "dV0", "dV1")}
    l, L = s.get("l"), s.get("L")
    tr_l, tr_c = s.get("tr_lambda"), s.get("tr_newcost")
    tr_a, tr_cl = s.get_int("tr_alpha"), s.get_int("tr_clamp")
    nbp = s.get_int("n_backpass")
    muf = s.get("mu_f") if s.L.lib.ilqgb_nx() and tr_l is not None else None
    recs = []
    for b in range(B):
        n = int(out["n_linesearch"][b])
        recs.append(dict(result=int(out["success"][b]), iterations=int(out["iterations"][b]), cost=out["cost"][b],
                         x=out["x"][b], u=out["u"][b], l=l[b], L=L[b], n_ls=n, n_bp=int(nbp[b]),
                         tr_lambda=tr_l[b][:n], tr_alpha=tr_a[b][:n], tr_newcost=tr_c[b][:n], tr_clamp=tr_cl[b],
                         **{k: v[b] for k, v in sc.items()}))
    s.close()
    return recs


def assert_same(gpu, ora, what, keys=("result", "iterations", "n_ls", "n_bp", "cost", "lambda", "g_norm", "w_pen_l",
                                       "w_pen_f", "tr_alpha", "tr_lambda", "tr_newcost", "x", "u")):
    for k in keys:
        a, b = np.asarray(gpu[k]), np.asarray(ora[k])
        assert a.shape == b.shape, f"{what}: {k} shape {a.shape} vs {b.shape}"
        assert np.array_equal(a, b), f"{what}: {k} differs (max abs diff {np.max(np.abs(a.astype(float) - b.astype(float)))})"
