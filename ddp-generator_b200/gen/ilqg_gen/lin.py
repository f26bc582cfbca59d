"""Expression linearisation: sympy expression -> explicit three-address fp64 statements.

Every operation becomes its own `const double tN = a op b;` statement, so the order of floating-point
operations is fixed by the generator, not by a compiler: the reference-ABI C file and the CUDA device
header print the SAME statement list and (with -ffp-contract=off / -fmad=false and dm_math.h) produce
bit-identical values.  Identical sub-expressions inside one scope are evaluated once (the reference relies
on its aux mechanism for that, gen_dep_graph.mac:186-229; here it is automatic).

Integer powers are expanded to multiplications, x**(1/2) to sqrt, negative powers to a division: no pow().
"""
from __future__ import annotations

import sympy as sp

_FUNCS = {
    sp.sin: "dm_sin",
    sp.cos: "dm_cos",
    sp.tan: "dm_tan",
    sp.asin: "dm_asin",
    sp.acos: "dm_acos",
}


def _lit(v):
    f = float(v)
    if f != f or f in (float("inf"), float("-inf")):
        raise ValueError("non-finite literal")
    r = repr(f)
    if "e" not in r and "." not in r and "n" not in r:
        r += ".0"
    return r


class Scope:
    """One straight-line scope; temps are shared by everything linearised into it."""

    def __init__(self, prefix="t"):
        self.prefix = prefix
        self.memo = {}
        self.sincos = {}
        self.items = []   # ("tmp", name, rhs_tokens) | ("out", target, operand, guard) | ("raw", payload)
        self.n = 0

    # operands: ("t", name) | ("lit", str) | ("sym", Symbol)
    def _new(self, rhs):
        name = f"{self.prefix}{self.n}"
        self.n += 1
        self.items.append(("tmp", name, rhs))
        return ("t", name)

    def ref(self, e):
        e = sp.sympify(e)
        if e in self.memo:
            return self.memo[e]
        r = self._build(e)
        self.memo[e] = r
        return r

    def _build(self, e):
        if e.is_Symbol:
            return ("sym", e)
        if e.is_Number:
            if e.is_negative:
                return ("lit", "-" + _lit(-e))
            return ("lit", _lit(e))
        if e is sp.S.true or e is sp.S.false:
            raise ValueError("boolean in arithmetic context")
        if e.is_Add:
            terms = e.as_ordered_terms()
            acc = self.ref(terms[0])
            for t in terms[1:]:
                if t.could_extract_minus_sign():
                    acc = self._new(("bin", "-", acc, self.ref(-t)))
                else:
                    acc = self._new(("bin", "+", acc, self.ref(t)))
            return acc
        if e.is_Mul:
            c, rest = e.as_coeff_Mul()
            if c != 1:
                if c == -1:
                    return self._new(("neg", self.ref(rest)))
                if c.is_negative:
                    return self._new(("neg", self.ref(-e)))
                return self._new(("bin", "*", self.ref(c), self.ref(rest)))
            num, den = [], []
            for fct in e.as_ordered_factors():
                if fct.is_Pow and fct.exp.is_Number and fct.exp.is_negative:
                    den.append(sp.Pow(fct.base, -fct.exp))
                else:
                    num.append(fct)
            nacc = None
            for fct in num:
                r = self.ref(fct)
                nacc = r if nacc is None else self._new(("bin", "*", nacc, r))
            if not den:
                return nacc
            dacc = None
            for fct in den:
                r = self.ref(fct)
                dacc = r if dacc is None else self._new(("bin", "*", dacc, r))
            if nacc is None:
                nacc = ("lit", "1.0")
            return self._new(("bin", "/", nacc, dacc))
        if e.is_Pow:
            b, ex = e.base, e.exp
            if not ex.is_Rational:
                raise ValueError(f"unsupported power {e}")
            if ex.is_negative:
                return self._new(("bin", "/", ("lit", "1.0"), self.ref(sp.Pow(b, -ex))))
            p, q = int(ex.p), int(ex.q)
            if q == 2:
                root = self._new(("call", "dm_sqrt", self.ref(b)))
                if p == 1:
                    return root
                return self._ipow(root, p)
            if q != 1:
                raise ValueError(f"unsupported power {e}")
            return self._ipow(self.ref(b), p)
        if isinstance(e, sp.Abs):
            return self._new(("call", "dm_fabs", self.ref(e.args[0])))
        if e.func in (sp.sin, sp.cos):
            # sin and cos of one argument come from ONE dm_sincos call (shared range reduction, straight-line code)
            arg = self.ref(e.args[0])
            if arg not in self.sincos:
                sn, cn = f"{self.prefix}{self.n}s", f"{self.prefix}{self.n}c"
                self.n += 1
                self.items.append(("sincos", sn, cn, arg))
                self.sincos[arg] = (("t", sn), ("t", cn))
            return self.sincos[arg][0 if e.func is sp.sin else 1]
        if e.func in _FUNCS:
            return self._new(("call", _FUNCS[e.func], self.ref(e.args[0])))
        if isinstance(e, sp.Piecewise):
            if len(e.args) != 2 or e.args[1][1] is not sp.S.true:
                raise ValueError(f"unsupported piecewise {e}")
            (ea, ca), (eb, _) = e.args
            cond = self._cond(ca)
            return self._new(("sel", cond, self.ref(ea), self.ref(eb)))
        raise ValueError(f"unsupported expression node {e.func}: {e}")

    def _ipow(self, base, p):
        # left-to-right repeated multiplication: x^3 = (x*x)*x
        acc = base
        for _ in range(p - 1):
            acc = self._new(("bin", "*", acc, base))
        return acc

    def _cond(self, c):
        ops = {sp.Ge: ">=", sp.Gt: ">", sp.Le: "<=", sp.Lt: "<"}
        for k, v in ops.items():
            if isinstance(c, k):
                return (v, self.ref(c.lhs), self.ref(c.rhs))
        raise ValueError(f"unsupported condition {c}")

    # outputs -------------------------------------------------------------------------------------------
    def out(self, target, e, guard):
        self.items.append(("out", target, self.ref(e), guard))

    def raw(self, payload):
        self.items.append(("raw", payload))


def render_operand(op, symname):
    kind = op[0]
    if kind == "t":
        return op[1]
    if kind == "lit":
        return op[1] if not op[1].startswith("-") else f"({op[1]})"
    return symname(op[1])


def render_rhs(rhs, symname):
    k = rhs[0]
    ro = lambda o: render_operand(o, symname)
    if k == "bin":
        return f"{ro(rhs[2])} {rhs[1]} {ro(rhs[3])}"
    if k == "neg":
        return f"-{ro(rhs[1])}"
    if k == "call":
        return f"{rhs[1]}({ro(rhs[2])})"
    if k == "sel":
        op, a, b = rhs[1]
        return f"({ro(a)} {op} {ro(b)}) ? {ro(rhs[2])} : {ro(rhs[3])}"
    raise ValueError(k)
