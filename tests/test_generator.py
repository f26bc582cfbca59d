"""The problem-code generator: committed outputs are reproducible, layouts follow the reference's templates, and the
generated derivatives agree with central finite differences of the generated dynamics/cost."""
import os
import tempfile

import numpy as np
import pytest
import sympy as sp

import oracle_lib
from ilqg_b200 import workloads as W
from ilqg_gen.__main__ import generate
from ilqg_gen.lower import lower, utri
from ilqg_gen.problems import REGISTRY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", sorted(REGISTRY))
def test_committed_code_is_reproducible(name):
    with tempfile.TemporaryDirectory() as d:
        generate(name, os.path.join(d, name))
        for f in ("iLQG_problem.h", "iLQG_func.c", f"{name}_device.cuh"):
            new = open(os.path.join(d, name, f)).read()
            old = open(os.path.join(ROOT, "ddp-generator_b200", "problems", name, f)).read()
            assert new == old, f"{name}/{f} is stale: run `python -m ilqg_gen` in ddp-generator_b200/gen"


def test_car_structure_facts():
    """Structure of the car problem as listed in SURVEY.md Appendix A.6."""
    m = lower(REGISTRY["car"]())
    assert (m.nx, m.nu, m.nqxx, m.nquu, m.nqxu) == (4, 2, 10, 3, 8)
    assert [d.name for d in m.params] == ["cf", "cu", "cx", "d", "h", "limA", "limW", "pf", "px"]
    assert sorted(e.idx for e in m.fx if e.time_var) == [8, 9, 12, 13, 14]        # (0,2),(1,2),(0,3),(1,3),(2,3)
    assert sorted(e.idx for e in m.fu if e.time_var) == [0, 1, 2]
    assert all(e.expr == 0 for e in m.cxu)
    assert sorted(e.idx for e in m.cxx if e.time_var) == [utri(0, 0), utri(1, 1)]
    assert [(h["input"], h["sign"]) for h in m.h] == [(0, -1), (0, 1), (1, -1), (1, 1)] and not m.has_hx
    assert sum(e.time_var for blk in (m.fxx, m.fuu, m.fxu) for e in blk) == 15


def test_brachi_constraint_folding():
    """hfe[1] becomes aux values hfe_1 / pfe_1 added to F, with the multiplier update of genenerator_main.mac:46-57."""
    m = lower(REGISTRY["brachi"]())
    assert [a.name for a in m.aux] == ["hfe_1", "pfe_1"] and m.n_mu == {"fe": 1, "fi": 0, "le": 0, "li": 0}
    mu, w_pen, h = m.mult["fe"][0]["mu"], m.w_pen, m.mult["fe"][0]["h"]
    assert sp.simplify(m.mult["fe"][0]["next"] - (mu + w_pen * h)) == 0
    pfe = [a for a in m.aux if a.name == "pfe_1"][0]
    assert sp.simplify(pfe.expr - (mu * h + sp.Rational(1, 2) * w_pen * h**2)) == 0
    d2 = [a for a in m.daux if a.name == "diff_2pfe_1_y_y"][0]
    assert m.Fcxx[0].expr == d2.sym and d2.expr == w_pen      # Fxx = w_pen


def _fd_check(problem, T, params, x0, u0, opts):
    """Derivative blocks stored by calc_derivs vs central differences of one-step rollouts of the same C code."""
    kind = "reference" if oracle_lib.available("reference", problem, 1) else "port"
    O = oracle_lib.OracleLib(kind, problem, 1)
    s = O.solver(T)
    s.set_opts(opts)
    s.set_params(params)
    assert s.init(x0, u0)
    s.set_scalar("w_pen_l", 1.0)
    s.set_scalar("w_pen_f", 1.0)
    assert s.calc_derivs()
    x, u = s.get("x"), s.get("u")
    fx, fu, cx, cu = s.get("fx"), s.get("fu"), s.get("cx"), s.get("cu")
    nx, nu = O.nx, O.nu
    one = O.solver(1)
    one.set_opts(opts)
    for name, size in zip(O.param_names, O.param_sizes):
        pass
    one.set_params({k: (v if O.param_sizes[O.param_names.index(k)] != -1 else v[:2]) for k, v in params.items()})

    def step(xk, uk):
        assert one.init(xk, uk[None])
        xs = one.get("x")
        return xs[1], one.get("c")[0]

    eps = 1e-6
    for k in (0, T // 3, T - 1):
        for j in range(nx):
            d = np.zeros(nx); d[j] = eps
            fp, cp = step(x[k] + d, u[k]); fm, cm = step(x[k] - d, u[k])
            assert np.allclose((fp - fm) / (2 * eps), fx[k].reshape(nx, nx, order="F")[:, j], rtol=1e-5, atol=1e-7)
            assert np.isclose((cp - cm) / (2 * eps), cx[k][j], rtol=1e-5, atol=1e-7)
        for j in range(nu):
            d = np.zeros(nu); d[j] = eps
            fp, cp = step(x[k], u[k] + d); fm, cm = step(x[k], u[k] - d)
            assert np.allclose((fp - fm) / (2 * eps), fu[k].reshape(nx, nu, order="F")[:, j], rtol=1e-5, atol=1e-7)
            assert np.isclose((cp - cm) / (2 * eps), cu[k][j], rtol=1e-5, atol=1e-7)


def test_car_derivatives_match_finite_differences():
    x0, u0 = W.car_single(T=60)
    u0 = np.clip(u0, -0.4, 0.4)     # stay inside the box so that clamping does not enter the differences
    _fd_check("car", 60, W.CAR_PARAMS, x0, u0, {"max_iter": 1})


def test_second_order_blocks_match_symbolic_hessians():
    """fxx/fuu/fxu/cxx/cuu entries of the generated C code vs sympy's own second derivatives evaluated numerically."""
    prob = REGISTRY["car"]()
    m = lower(prob)
    O = oracle_lib.OracleLib("port", "car", 1)
    T = 20
    s = O.solver(T)
    s.set_params(W.CAR_PARAMS)
    x0, u0 = W.car_single(T=T)
    u0 = np.clip(u0, -0.4, 0.4)
    assert s.init(x0, u0) and s.calc_derivs()
    x, u, fxx, fuu, fxu, cxx = s.get("x"), s.get("u"), s.get("fxx"), s.get("fuu"), s.get("fxu"), s.get("cxx")
    subs_p = {}
    for d in m.params:
        for sym, val in zip(d.symbols, W.CAR_PARAMS[d.name]):
            subs_p[sym] = val
    aux = {a.handle: a.definition for a in prob.aux}
    for k in (0, 7, 19):
        pt = {**subs_p, **dict(zip(m.x, x[k])), **dict(zip(m.u, u[k]))}
        f = [sp.sympify(prob.f[xs]).subs(aux) for xs in m.x]
        for i in range(4):
            for a in range(4):
                for b in range(a, 4):
                    want = float(sp.diff(f[i], m.x[a], m.x[b]).subs(pt))
                    assert np.isclose(fxx[k][i * 10 + utri(a, b)], want, rtol=1e-9, atol=1e-12)
            for a in range(2):
                for b in range(a, 2):
                    want = float(sp.diff(f[i], m.u[a], m.u[b]).subs(pt))
                    assert np.isclose(fuu[k][i * 3 + utri(a, b)], want, rtol=1e-9, atol=1e-12)
            for a in range(4):
                for b in range(2):
                    want = float(sp.diff(f[i], m.x[a], m.u[b]).subs(pt))
                    assert np.isclose(fxu[k][i * 8 + a + b * 4], want, rtol=1e-9, atol=1e-12)
        Lc = sp.sympify(prob.L)
        for a in range(4):
            for b in range(a, 4):
                assert np.isclose(cxx[k][utri(a, b)], float(sp.diff(Lc, m.x[a], m.x[b]).subs(pt)), rtol=1e-9, atol=1e-12)
