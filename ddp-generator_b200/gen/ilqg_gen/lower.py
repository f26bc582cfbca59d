"""Lowering: problem description -> derivative blocks with the reference's layouts and semantics.

Restates what the reference's Maxima generator does (citations into /root/reference):
  * augmented-Lagrangian folding of hfe/hfi/hle/hli into F and L as auxiliary penalty values
    (genenerator_main.mac:46-124),
  * auxiliary values and their first/second derivatives as named, separately evaluated quantities with
    chain-rule references between them (gen_dep_graph.mac:186-229),
  * "time-varying" classification -- anything depending on x, u, w_pen, a multiplier or a [k]-indexed
    parameter is recomputed every pass, the rest is written once (gen_dep_graph.mac:143-171;
    iLQG_func.tem:312-362),
  * derivative arrays fx, fu, fxx, fuu, fxu, Fx, Fxx, Lx, Lxx, Lu, Luu, Lxu and their packed layouts
    (genenerator_main.mac:215-371; matMult.h:4-9),
  * input constraints h[i] -> clamp / shifted limits / active-constraint gradient
    (genenerator_main.mac:373-447).
The emitters (emit_c.py / emit_cuda.py) only print what this module computed.
"""
from __future__ import annotations

import sympy as sp

from .problem import Problem


def utri(r, c):
    """Packed upper-triangle index, column by column (matMult.h:8)."""
    assert r <= c
    return (c * (c + 1)) // 2 + r


class Entry:
    __slots__ = ("idx", "expr", "time_var", "atomic")

    def __init__(self, idx, expr, time_var):
        self.idx = idx
        self.expr = expr
        self.time_var = time_var
        self.atomic = expr.is_Number or expr.is_Symbol


class AuxInfo:
    def __init__(self, name, sym, expr, time_var, dep_u):
        self.name = name
        self.sym = sym
        self.expr = expr
        self.time_var = time_var
        self.dep_u = dep_u
        self.used_running = False
        self.used_final = False
        self.atomic = expr.is_Number or expr.is_Symbol


class Model:
    pass


def lower(prob: Problem, full_ddp_blocks=True) -> Model:
    prob.validate()
    m = Model()
    m.name = prob.name
    m.x = list(prob.x)
    m.u = list(prob.u)
    nx, nu = len(m.x), len(m.u)
    m.nx, m.nu = nx, nu
    xu = m.x + m.u
    xu_pos = {s: i for i, s in enumerate(xu)}

    w_pen = sp.Symbol("w_pen", real=True)
    m.w_pen = w_pen
    half = sp.Rational(1, 2)

    # ---- constraint folding (genenerator_main.mac:46-124) ---------------------------------------------
    F = sp.sympify(prob.F)
    L = sp.sympify(prob.L)
    m.mult = {"fe": [], "fi": [], "le": [], "li": []}
    mu_syms = []

    def fold(kind, exprs):
        nonlocal F, L
        for i, hexpr in enumerate(exprs, start=1):
            mu = sp.Symbol(f"mu_{kind}_{i}", real=True)
            mu_syms.append(mu)
            hh = prob.def_aux(f"h{kind}_{i}", sp.sympify(hexpr))
            rec = {"i": i - 1, "mu": mu, "h_handle": hh}
            if kind in ("fe", "le"):
                pen = mu * hh + half * w_pen * hh**2
                rec["next"] = mu + w_pen * hh
            else:
                pen = sp.Piecewise((mu * hh * (1 + w_pen * hh), hh >= 0), (mu * hh / (1 - w_pen * hh), True))
                rec["next_A"] = mu * (1 + 2 * w_pen * hh)
                rec["next_I"] = mu * (1 - w_pen * hh) ** -2
            ph = prob.def_aux(f"p{kind}_{i}", pen)
            if kind in ("fe", "fi"):
                F = F + ph
            else:
                L = L + ph
            m.mult[kind].append(rec)

    fold("fe", prob.hfe)
    fold("fi", prob.hfi)
    fold("le", prob.hle)
    fold("li", prob.hli)
    m.mu_syms = mu_syms

    # ---- parameters, sorted by name (genenerator_main.mac:164-169) -----------------------------------
    m.params = [prob.params[k] for k in sorted(prob.params.keys())]
    m.param_leaf = {}
    flat = 0
    for pi, d in enumerate(m.params):
        d.index = pi
        d.flat_offset = flat
        if d.size == -1:
            m.param_leaf[d.symbols[0]] = (pi, "k")
            flat += 1   # flat offset is only meaningful for time-invariant parameters
        else:
            for ei, s in enumerate(d.symbols):
                m.param_leaf[s] = (pi, ei)
            flat += d.size
    m.n_param_flat = flat
    m.has_k_params = any(d.size == -1 for d in m.params)
    kparam_syms = {d.symbols[0] for d in m.params if d.size == -1}

    # ---- aux symbols & lowering of expressions --------------------------------------------------------
    aux_by_handle = {}
    aux_sym = {}
    for a in prob.aux:
        s = sp.Symbol(f"aux_{a.name}", real=True)
        aux_by_handle[a.handle] = a
        aux_sym[a.name] = s
    handle_to_sym = {a.handle: aux_sym[a.name] for a in prob.aux}

    tv_syms = set(xu) | {w_pen} | set(mu_syms) | kparam_syms
    m.aux = []       # AuxInfo in order
    m.daux = []      # AuxInfo for derivatives, creation order re-sorted below
    daux_memo = {}

    def is_time_var(e):
        return bool(e.free_symbols & tv_syms)

    def get_daux(a, vars_):
        vars_ = tuple(sorted(vars_, key=lambda s: xu_pos[s]))
        key = (a.name, vars_)
        if key in daux_memo:
            return daux_memo[key]
        raw = sp.diff(a.definition, *vars_)
        low = lower_expr(raw)
        if low.is_Number:
            daux_memo[key] = low
            return low
        if len(vars_) == 1:
            dname = f"diff_{a.name}_{vars_[0].name}"
        else:
            dname = f"diff_2{a.name}_{vars_[0].name}_{vars_[1].name}"
        s = sp.Symbol(f"daux_{dname}", real=True)
        info = AuxInfo(dname, s, low, is_time_var(low), any(v in m.u for v in a.deps))
        info.aux_index = [x.name for x in prob.aux].index(a.name)
        info.order = len(vars_)
        info.var_pos = tuple(xu_pos[v] for v in vars_)
        if info.time_var:
            tv_syms.add(s)
        m.daux.append(info)
        daux_memo[key] = s
        return s

    def lower_expr(e):
        e = sp.sympify(e)
        rep = {}
        for d in e.atoms(sp.Derivative):
            h = d.expr
            if h not in aux_by_handle:
                raise ValueError(f"cannot differentiate {d}")
            if len(d.variables) > 2:
                raise ValueError(f"third derivative of aux requested: {d}")
            rep[d] = get_daux(aux_by_handle[h], d.variables)
        if rep:
            e = e.xreplace(rep)
        e = e.xreplace(handle_to_sym)
        return e

    for a in prob.aux:
        low = lower_expr(a.definition)
        info = AuxInfo(a.name, aux_sym[a.name], low, is_time_var(low), any(v in m.u for v in a.deps))
        if info.time_var:
            tv_syms.add(info.sym)
        m.aux.append(info)

    def D(e, *vars_):
        return lower_expr(sp.diff(e, *vars_))

    # ---- dynamics, costs --------------------------------------------------------------------------------
    f_raw = [sp.sympify(prob.f[s]) for s in m.x]
    m.f = [lower_expr(e) for e in f_raw]
    m.L = lower_expr(L)
    m.F = lower_expr(F)

    # ---- derivative blocks (genenerator_main.mac:333-371; layouts 215-329) ------------------------------
    def entry(idx, e):
        return Entry(idx, e, is_time_var(e))

    m.fx = [entry(r + c * nx, D(f_raw[r], m.x[c])) for c in range(nx) for r in range(nx)]
    m.fu = [entry(r + c * nx, D(f_raw[r], m.u[c])) for c in range(nu) for r in range(nx)]
    nqxx, nquu, nqxu = nx * (nx + 1) // 2, nu * (nu + 1) // 2, nx * nu
    m.nqxx, m.nquu, m.nqxu = nqxx, nquu, nqxu
    m.fxx = [entry(i * nqxx + utri(j, k), D(f_raw[i], m.x[j], m.x[k]))
             for i in range(nx) for k in range(nx) for j in range(k + 1)]
    m.fuu = [entry(i * nquu + utri(j, k), D(f_raw[i], m.u[j], m.u[k]))
             for i in range(nx) for k in range(nu) for j in range(k + 1)]
    m.fxu = [entry(i * nqxu + j + k * nx, D(f_raw[i], m.x[j], m.u[k]))
             for i in range(nx) for k in range(nu) for j in range(nx)]
    m.cx = [entry(i, D(L, m.x[i])) for i in range(nx)]
    m.cxx = [entry(utri(r, c), D(L, m.x[r], m.x[c])) for c in range(nx) for r in range(c + 1)]
    m.cu = [entry(i, D(L, m.u[i])) for i in range(nu)]
    m.cuu = [entry(utri(r, c), D(L, m.u[r], m.u[c])) for c in range(nu) for r in range(c + 1)]
    m.cxu = [entry(i + j * nx, D(L, m.x[i], m.u[j])) for j in range(nu) for i in range(nx)]
    m.Fcx = [entry(i, D(F, m.x[i])) for i in range(nx)]
    m.Fcxx = [entry(utri(r, c), D(F, m.x[r], m.x[c])) for c in range(nx) for r in range(c + 1)]

    # ---- user outputs g (iLQG_func.tem:511-521): plain expressions of x, u, parameters and auxiliary values ----------------
    m.g = [lower_expr(e) for e in getattr(prob, "g", [])]

    # ---- input constraints (genenerator_main.mac:373-447) ---------------------------------------------
    m.h = []
    for hi, hexpr in enumerate(prob.h):
        hexpr = sp.sympify(hexpr)
        hu = [sp.diff(hexpr, uu) for uu in m.u]
        nz = [j for j, d in enumerate(hu) if d != 0]
        if len(nz) != 1:
            raise ValueError(f"constraint ({hexpr}) may depend on only one input")
        j = nz[0]
        sign = hu[j]
        if sign not in (1, -1):
            raise ValueError(f"coefficient of input in constraint ({hexpr}) must be 1 or -1")
        # remaining dependence on u (through aux) is not allowed
        lim = sp.expand(hexpr - sign * m.u[j])
        if any(uu in lim.free_symbols for uu in m.u):
            raise ValueError(f"constraint ({hexpr}) may only depend directly on one input")
        if sign > 0:
            lim = -lim
        rec = {
            "index": hi,
            "input": j,
            "sign": int(sign),            # +1: upper bound, -1: lower bound
            "limit": lower_expr(lim),
            "hx": [D(hexpr, xx) for xx in m.x],
        }
        m.h.append(rec)
    m.has_hx = any(e != 0 for rec in m.h for e in rec["hx"])

    # ---- multiplier updates (iLQG_func.tem:417-509) ------------------------------------------------------
    for kind in ("fe", "fi", "le", "li"):
        for rec in m.mult[kind]:
            rec["h"] = lower_expr(rec["h_handle"])
            for key in ("next", "next_A", "next_I"):
                if key in rec:
                    rec[key] = lower_expr(rec[key])

    # ---- ordering of aux derivatives: by owning aux, then order, then variable positions -----------------
    m.daux.sort(key=lambda d: (d.aux_index, d.order, d.var_pos))

    # ---- used_by_running / used_by_final closure (gen_dep_graph.mac:143-147) -----------------------------
    by_sym = {a.sym: a for a in m.aux}
    by_sym.update({d.sym: d for d in m.daux})

    def mark(exprs, attr):
        stack = []
        for e in exprs:
            stack.extend(s for s in sp.sympify(e).free_symbols if s in by_sym)
        while stack:
            s = stack.pop()
            info = by_sym[s]
            if getattr(info, attr):
                continue
            setattr(info, attr, True)
            stack.extend(t for t in info.expr.free_symbols if t in by_sym)

    running_roots = list(m.f) + [m.L]
    for blk in (m.fx, m.fu, m.cx, m.cxx, m.cu, m.cuu, m.cxu):
        running_roots += [e.expr for e in blk]
    if full_ddp_blocks:
        for blk in (m.fxx, m.fuu, m.fxu):
            running_roots += [e.expr for e in blk]
    for rec in m.h:
        running_roots += [rec["limit"]] + rec["hx"]
    running_roots += list(m.g)
    for kind in ("le", "li"):
        for rec in m.mult[kind]:
            running_roots += [rec["h"]] + [rec[k] for k in ("next", "next_A", "next_I") if k in rec]
    final_roots = [m.F] + [e.expr for e in m.Fcx] + [e.expr for e in m.Fcxx]
    for kind in ("fe", "fi"):
        for rec in m.mult[kind]:
            final_roots += [rec["h"]] + [rec[k] for k in ("next", "next_A", "next_I") if k in rec]
    mark(running_roots, "used_running")
    mark(final_roots, "used_final")
    m.tv_syms = tv_syms
    m.uses_w_pen_running = any(w_pen in sp.sympify(e).free_symbols for e in running_roots) or any(
        w_pen in a.expr.free_symbols for a in m.aux + m.daux if a.used_running)
    m.n_mu = {k: len(v) for k, v in m.mult.items()}
    return m
