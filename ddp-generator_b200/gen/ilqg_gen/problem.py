"""Problem-description front end (Python equivalent of the reference's Maxima `.mac` problem files).

A problem file of the reference defines `x`, `u`, `f[...]`, `L`, `F`, optional `h[i]` (input box constraints),
`hfe/hfi` (terminal equality / inequality), `hle/hli` (running equality / inequality) and auxiliary values
written with an apostrophe (README.md:25-38 of the reference; checks in genenerator_main.mac:2-27).  This
module offers the same vocabulary as a small sympy-based DSL; `lower.py` turns it into derivative blocks.
"""
from __future__ import annotations

import sympy as sp


class AuxVar:
    """Auxiliary value: evaluated once per timestep, reused by f/L/F/h and their derivatives
    (reference: README.md:38, gen_dep_graph.mac:186-229)."""

    def __init__(self, name, definition, deps, handle):
        self.name = name
        self.definition = definition  # sympy expr in states/inputs/params/other aux handles
        self.deps = deps              # ordered list of state/input symbols it depends on (transitively)
        self.handle = handle          # applied sympy Function (or Symbol if no x/u dependence)


class ParamDesc:
    def __init__(self, name, size):
        self.name = name
        self.size = size      # 1 scalar, k>1 fixed array, -1 => indexed by timestep k (length n_hor+1)
        self.symbols = []     # sympy symbols for the elements (one symbol for scalar / [k] kinds)


class Problem:
    def __init__(self, name):
        self.name = name
        self.x = []
        self.u = []
        self.params = {}      # name -> ParamDesc
        self.aux = []         # AuxVar, in definition (= topological) order
        self.f = {}           # state symbol -> next-state expression
        self.L = sp.Integer(0)
        self.F = sp.Integer(0)
        self.h = []           # input constraints h[i] < 0 (each depends on exactly one input with coeff +-1)
        self.hfe = []
        self.hfi = []
        self.hle = []
        self.hli = []
        self.g = []           # optional user outputs g[i](x, u, p) evaluated by calcG (iLQG_func.tem:511-521)

    # --- symbols -------------------------------------------------------------------------------------
    def states(self, names):
        syms = list(sp.symbols(names, real=True, seq=True))
        self.x.extend(syms)
        return syms

    def inputs(self, names):
        syms = list(sp.symbols(names, real=True, seq=True))
        self.u.extend(syms)
        return syms

    def param(self, name):
        d = ParamDesc(name, 1)
        d.symbols = [sp.Symbol(name, real=True)]
        self.params[name] = d
        return d.symbols[0]

    def param_array(self, name, size):
        d = ParamDesc(name, size)
        d.symbols = [sp.Symbol(f"{name}_{i}", real=True) for i in range(size)]
        self.params[name] = d
        return d.symbols

    def param_k(self, name):
        """Parameter indexed by the timestep, `name[k]` in the reference (genenerator_main.mac:131-162)."""
        d = ParamDesc(name, -1)
        d.symbols = [sp.Symbol(f"{name}_k", real=True)]
        self.params[name] = d
        return d.symbols[0]

    def def_aux(self, name, expr):
        expr = sp.sympify(expr)
        xu = self.x + self.u
        deps = [s for s in xu if s in expr.free_symbols]
        if deps:
            handle = sp.Function(name, real=True)(*deps)
        else:
            handle = sp.Symbol(f"auxsym_{name}", real=True)
        a = AuxVar(name, expr, deps, handle)
        self.aux.append(a)
        return handle

    # --- validation (genenerator_main.mac:2-27, 373-395) -----------------------------------------------
    def validate(self):
        if not self.x:
            raise ValueError("vector of states x not defined")
        if not self.u:
            raise ValueError("vector of inputs u not defined")
        if set(self.f.keys()) != set(self.x):
            raise ValueError("elements of f must be indexed by elements of x")
        for name, lst in (("hfe", self.hfe), ("hfi", self.hfi)):
            for e in lst:
                if any(s in sp.sympify(e).free_symbols for s in self.u):
                    raise ValueError(f"{name} must not depend on any input u")
        if any(s in sp.sympify(self.F).free_symbols for s in self.u):
            raise ValueError("F may not depend on u")
