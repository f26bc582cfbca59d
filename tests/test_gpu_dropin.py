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


DERIV_FIELDS = ("fx", "fu", "cx", "cxx", "cu", "cuu", "cxu", "lower", "upper", "lower_sign", "upper_sign", "lower_hx", "upper_hx")


def _phase_loop(kind, problem, ddp, T, params, x0, u0, opts, n_iter):
    """A replica of the iLQG() loop (iLQG.c:239-363) built from the solver's PUBLIC phases -- calc_derivs, back_pass,
    line_search, update_multipliers, forward_pass, makeCandidateNominal -- with a snapshot after every phase."""
    s = oracle_lib.OracleLib(kind, problem, ddp).solver(T)
    s.set_opts(opts)
    s.set_params(params)
    assert s.init(x0, u0)
    s.set_scalar("lambda", 1.0); s.set_scalar("w_pen_l", 1.0); s.set_scalar("w_pen_f", 1.0)
    s.update_multipliers(1)
    snaps = [("init", {"mult_t": s.get("mult_t"), "mult_f": s.get("mult_f")})]
    lam, dl, new_deriv = 1.0, 1.0, True
    f2 = opts.get("w_pen_fact2", 1.0)
    for it in range(n_iter):
        if new_deriv:
            assert s.calc_derivs()
            snaps.append((f"derivs{it}", {f: s.get(f) for f in DERIV_FIELDS}))
            new_deriv = False
        fails = 0
        while s.back_pass():
            fails += 1
            snaps.append((f"bp_fail{it}.{fails}", {"l": s.get("l"), "L": s.get("L"), "dV0": s.scalar("dV0"), "dV1": s.scalar("dV1")}))
            dl = max(dl * 1.6, 1.6); lam = max(lam * dl, 1e-6); s.set_scalar("lambda", lam)
            assert fails < 60
        snaps.append((f"bp{it}", {"l": s.get("l"), "L": s.get("L"), "fails": fails, **{k: s.scalar(k) for k in ("dV0", "dV1", "g_norm")}}))
        ok = s.line_search(it)
        snaps.append((f"ls{it}", {"ok": ok, "log": s.get("log_linesearch")[it], **{k: s.scalar(k) for k in ("new_cost", "dcost", "expected")}}))
        if ok:
            dl = min(dl / 1.6, 1 / 1.6); lam = lam * dl * (lam > 1e-6); s.set_scalar("lambda", lam)
            s.make_candidate_nominal(); s.set_scalar("cost", s.scalar("new_cost")); new_deriv = True
            snaps.append((f"accept{it}", {"x": s.get("x"), "u": s.get("u")}))
            if s.scalar("dcost") < 1e-7:
                break
            s.update_multipliers(0)
            s.set_scalar("cost", s.forward_pass(0.0, cost_only=1)[1])
            snaps.append((f"mult{it}", {"mult_t": s.get("mult_t"), "mult_f": s.get("mult_f"), "cost": s.scalar("cost"),
                                        "w_pen_l": s.scalar("w_pen_l"), "w_pen_f": s.scalar("w_pen_f")}))
        else:
            dl = max(dl * 1.6, 1.6); lam = max(lam * dl, 1e-6); s.set_scalar("lambda", lam)
            if f2 > 1.0:
                s.set_scalar("w_pen_l", s.scalar("w_pen_l") * f2); s.set_scalar("w_pen_f", s.scalar("w_pen_f") * f2)
                s.set_scalar("cost", s.forward_pass(0.0, cost_only=1)[1])
    s.close()
    return snaps


@pytest.mark.parametrize("problem,ddp,n_iter", [("car", 0, 8), ("car", 1, 4), ("carhx", 0, 5), ("brachi_hli", 0, 6), ("pend", 0, 8), ("pend", 1, 6)])
def test_public_phases_through_reference_api(problem, ddp, n_iter):
    """calc_derivs / back_pass / line_search / update_multipliers of the GPU drop-in against the reference's, phase by phase,
    through the same harness: every member each function writes must be bit-identical (boundary row)."""
    if problem in ("car", "carhx"):
        T, (x0, u0), opts = 200, W.car_batch(1, T=200, seed=21), {"max_iter": 20}
        x0, u0, params = x0[0], u0[0], (W.CAR_PARAMS if problem == "car" else W.CARHX_PARAMS)
    elif problem == "pend":
        xb, ub = W.pend_batch(4)
        T, x0, u0, params, opts = W.PEND_T, xb[1], ub[1], W.PEND_PARAMS, dict(W.PEND_OPTS)
    else:
        params, x0, u0, opts = W.brachi_hli(100)
        T = 100
    kind = PU.oracle_kinds(problem, ddp)[0]
    a = _phase_loop(kind, problem, ddp, T, params, x0, u0, opts, n_iter)
    b = _phase_loop("b200", problem, ddp, T, params, x0, u0, opts, n_iter)
    assert [n for n, _ in a] == [n for n, _ in b]
    for (name, ra), (_, rb) in zip(a, b):
        for k in ra:
            assert np.array_equal(np.asarray(ra[k]), np.asarray(rb[k])), f"{problem} ddp{ddp}: {k} differs after {name}"


def test_clampU_through_reference_api():
    """clampU(u, t, k, p, N) (iLQG_func.tem:68) incl. a state-dependent limit (carhx: the steering limit shrinks with speed)"""
    import ctypes as C
    T = 40
    x0, u0 = W.car_batch(1, T=T, seed=23)
    for problem, params in (("car", W.CAR_PARAMS), ("carhx", W.CARHX_PARAMS)):
        outs = []
        for kind in (PU.oracle_kinds(problem, 0)[0], "b200"):
            O = oracle_lib.OracleLib(kind, problem, 0)
            O.lib.h_clamp_u.argtypes = [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(dtype=np.float64)]
            O.lib.h_clamp_u.restype = None
            s = O.solver(T)
            s.set_params(params)
            assert s.init(x0[0], u0[0])
            res = []
            for k in (0, 7, 39):
                for u in ([2.0, -5.0], [-0.9, 3.5], [0.1, 0.2], [-0.45, 1.9]):
                    ua = np.array(u)
                    O.lib.h_clamp_u(s.h, k, ua)
                    res.append(ua.copy())
            outs.append(np.stack(res))
            s.close()
        assert np.array_equal(outs[0], outs[1]), problem
        assert not np.array_equal(outs[0][0], [2.0, -5.0])


def test_user_outputs_through_reference_api():
    """get_g_size / calcG (iLQG.h:87-88, iLQG_func.tem:511-521): the pend problem defines two outputs g[]"""
    import ctypes as C
    x0, u0 = W.pend_batch(1)
    outs = []
    for kind in (PU.oracle_kinds("pend", 0)[0], "b200"):
        O = oracle_lib.OracleLib(kind, "pend", 0)
        O.lib.h_calc_g.argtypes = [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(dtype=np.float64)]
        s = O.solver(W.PEND_T)
        s.set_opts(W.PEND_OPTS); s.set_params(W.PEND_PARAMS)
        assert s.init(x0[0], u0[0])
        res = []
        for k in (0, 17, W.PEND_T - 1):
            g = np.zeros(4)
            assert O.lib.h_calc_g(s.h, k, g) == 2
            res.append(g[:2].copy())
        outs.append(np.stack(res))
        s.close()
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).min() > 0
    c = oracle_lib.OracleLib("b200", "car", 0)
    c.lib.h_calc_g.argtypes = [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(dtype=np.float64)]
    sc = c.solver(10); sc.set_params(W.CAR_PARAMS)
    xc, uc = W.car_batch(1, T=10)
    assert sc.init(xc[0], uc[0]) and c.lib.h_calc_g(sc.h, 3, np.zeros(2)) == 0       # a problem without outputs: size 0
    sc.close()
