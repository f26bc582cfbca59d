"""Phase-level GPU parity: each kernel on its own against the matching public function of the reference
(calc_derivs, back_pass, line_search -- iLQG.h:83, back_pass.h:7, line_search.h:6), field by field, bit-exact."""
import numpy as np
import pytest

import ilqg_b200
import oracle_lib
import parity_util as PU
from ilqg_b200 import workloads as W
from ilqg_gen.lower import lower
from ilqg_gen.problems import REGISTRY

pytestmark = pytest.mark.gpu


def varying_map(problem, full):
    """(field, index) of every entry of the device's time-varying derivative record, in its canonical order."""
    m = lower(REGISTRY[problem]())
    blocks = [("fx", m.fx), ("fu", m.fu), ("cx", m.cx), ("cxx", m.cxx), ("cu", m.cu), ("cuu", m.cuu), ("cxu", m.cxu)]
    v1 = [(k, e.idx) for k, es in blocks for e in es if e.time_var]
    v1 += [("lower", i) for i in range(m.nu)] + [("upper", i) for i in range(m.nu)]
    v2 = [(k, e.idx) for k, es in (("fxx", m.fxx), ("fuu", m.fuu), ("fxu", m.fxu)) for e in es if e.time_var]
    return v1, (v2 if full else [])


@pytest.mark.parametrize("problem,ddp,T", [("car", 0, 500), ("car", 1, 120), ("brachi", 0, 50), ("brachi", 1, 5)])
def test_phases_match_oracle(problem, ddp, T):
    B = 5
    if problem == "car":
        x0, u0 = W.car_batch(B, T=T, seed=11)
        params, opts = W.CAR_PARAMS, {"max_iter": 10}
    else:
        params, bx0, bu0, opts = W.brachi(T)
        x0 = np.repeat(bx0[None], B, 0) * (1 + np.arange(B))[:, None]
        u0 = np.repeat(bu0[None], B, 0) * (1 + 0.1 * np.arange(B))[:, None, None]
    kind = PU.oracle_kinds(problem, ddp)[0]
    O = oracle_lib.OracleLib(kind, problem, ddp)
    g = ilqg_b200.BatchSolver(problem, ddp, B, T, flags=ilqg_b200.TRACE)
    g.set_options(opts); g.set_params(params); g.upload(x0, u0); g.start()
    v1map, v2map = varying_map(problem, ddp)
    sols = []
    for b in range(B):
        s = O.solver(T); s.set_opts(opts); s.set_params(params)
        assert s.init(x0[b], u0[b])
        # iLQG() would set these before the first pass (iLQG.c:226-237)
        s.set_scalar("lambda", 1.0); s.set_scalar("w_pen_l", 1.0); s.set_scalar("w_pen_f", 1.0)
        s.update_multipliers(1)
        sols.append(s)
    assert np.array_equal(g.get("cost"), [s.scalar("cost") for s in sols])
    assert np.array_equal(g.get("x"), np.stack([s.get("x") for s in sols]))
    need_derivs = [True] * B      # newDeriv of iLQG.c:241: derivatives are only refreshed after an accepted step
    for it in range(3):
        # ---- derivative pass ----
        g.phase("derivs")
        v1 = g.get("v1").reshape(B, T, -1)
        v2 = g.get("v2").reshape(B, T, -1) if ddp else None
        fd = g.get("fd")
        for b, s in enumerate(sols):
            if need_derivs[b]:
                assert s.calc_derivs()
            fields = {}
            for j, (f, idx) in enumerate(v1map):
                fields.setdefault(f, s.get(f))
                assert np.array_equal(v1[b, :, j], fields[f][:T, idx]), (it, b, f, idx)
            for j, (f, idx) in enumerate(v2map):
                fields.setdefault(f, s.get(f))
                assert np.array_equal(v2[b, :, j], fields[f][:T, idx]), (it, b, f, idx)
            assert np.array_equal(fd[b], np.concatenate([s.get("cx")[T], s.get("cxx")[T]]))
        # ---- backward pass (the device kernel also retries with larger lambda like iLQG.c:261-284) ----
        g.phase("backpass")
        for b, s in enumerate(sols):
            lam = s.scalar("lambda")
            dl = 1.0
            while s.back_pass():
                dl = max(dl * 1.6, 1.6); lam = max(lam * dl, 1e-6); s.set_scalar("lambda", lam)
        # problems whose lambda schedule already moved are compared through the values the line search needs
        for k in ("dV0", "dV1", "g_norm"):
            assert np.array_equal(g.get(k), [s.scalar(k) for s in sols]), (it, k)
        assert np.array_equal(g.get("l"), np.stack([s.get("l") for s in sols]))
        assert np.array_equal(g.get("L"), np.stack([s.get("L") for s in sols]))
        # ---- line search + accept ----
        g.phase("linesearch")
        for b, s in enumerate(sols):
            ok = s.line_search(it)
            need_derivs[b] = bool(ok)
            assert g.get("new_cost")[b] == s.scalar("new_cost")
            if ok:      # iLQG.c:311-338
                s.make_candidate_nominal(); s.set_scalar("cost", s.scalar("new_cost"))
                if s.scalar("dcost") >= 1e-7:
                    s.update_multipliers(0)
                    s.set_scalar("cost", s.forward_pass(0.0, cost_only=1)[1])
            elif opts.get("w_pen_fact2", 1.0) > 1.0:      # iLQG.c:345-349
                s.set_scalar("w_pen_l", s.scalar("w_pen_l") * opts["w_pen_fact2"])
                s.set_scalar("w_pen_f", s.scalar("w_pen_f") * opts["w_pen_fact2"])
                s.set_scalar("cost", s.forward_pass(0.0, cost_only=1)[1])
        assert np.array_equal(g.get("x"), np.stack([s.get("x") for s in sols])), it
        assert np.array_equal(g.get("u"), np.stack([s.get("u") for s in sols])), it
        assert np.array_equal(g.get("cost"), [s.scalar("cost") for s in sols]), it
        assert np.array_equal(g.get("w_pen_f"), [s.scalar("w_pen_f") for s in sols]), it
        # keep lambda in step with the device for the next pass
        lam_dev = g.get("lambda")
        for b, s in enumerate(sols):
            s.set_scalar("lambda", lam_dev[b])
        if not (g.get_int("status") == 0).all():
            break
    g.close()
