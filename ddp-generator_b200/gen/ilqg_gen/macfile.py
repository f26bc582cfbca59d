"""Front end for the reference's Maxima problem files (`.mac`): the subset its examples use, so that an existing
problem description works unchanged (format: README.md:25-38, examples/*/optDef*.mac).

Understood statements (terminated by `;` or `$`, `/* ... */` comments ignored):
    x: [s1, s2, ...];  u: [i1, ...];              state / input lists
    name: expr;                                     a value; referenced as 'name it is an auxiliary value (evaluated
                                                    once per step and shared), referenced as name it is substituted
    f[state]: expr;  L: expr;  F: expr;             dynamics, running cost, final cost
    h[i]: expr;  hfe[i] / hfi[i] / hle[i] / hli[i]  input box constraints, terminal / running (in)equalities
    g[i]: expr;                                     user outputs evaluated by calcG (iLQG_func.tem:511-521)
    fname(a, b, ...):= expr;                        helper function
    assume(sym > 0);  assume(sym < 0);              sign assumptions (used by abs / integrate)
Expressions: + - * / ^, numbers, sqrt sin cos tan asin acos abs, integrate(e, v, a, b), expand, factor, helper calls.
Every other undefined identifier is a parameter: `p` a scalar, `p[3]` an array element, `p[k]` one value per timestep.
"""
from __future__ import annotations

import ast
import re

import sympy as sp

from .problem import Problem

_FUNCS = {"sqrt": sp.sqrt, "sin": sp.sin, "cos": sp.cos, "tan": sp.tan, "asin": sp.asin, "acos": sp.acos, "abs": sp.Abs,
          "expand": sp.expand, "factor": sp.factor, "ratsimp": sp.ratsimp}
_SPECIAL_ARRAYS = ("f", "h", "hfe", "hfi", "hle", "hli", "g")


class _ParamArray:
    def __init__(self, name, assumptions):
        self.name = name
        self.elems = {}
        self.k_sym = None
        self.assumptions = assumptions

    def __getitem__(self, idx):
        if isinstance(idx, (int, sp.Integer)):
            i = int(idx)
            if self.k_sym is not None:
                raise ValueError(f"found index k and other integer index mixed in param {self.name}")
            return self.elems.setdefault(i, sp.Symbol(f"{self.name}_{i}", real=True, **self.assumptions))
        if isinstance(idx, sp.Symbol) and idx.name == "k":
            if self.elems:
                raise ValueError(f"found index k and other integer index mixed in param {self.name}")
            if self.k_sym is None:
                self.k_sym = sp.Symbol(f"{self.name}_k", real=True, **self.assumptions)
            return self.k_sym
        raise ValueError(f"only integer indices or k supported, found {idx} in param {self.name}")


def _statements(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return [s.strip() for s in re.split(r"[;$]", text) if s.strip()]


_AST_OK = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.Call, ast.Name, ast.Constant, ast.Subscript, ast.Load, ast.Add, ast.Sub,
           ast.Mult, ast.Div, ast.Pow, ast.USub, ast.UAdd)


class _ExactInts(ast.NodeTransformer):
    """Integer literals become exact sympy Integers, as in Maxima: 2/3 is a rational and p^(1/2) a square root, not Python's
    float division.  Subscripts (cf[2], ymin[k]) keep plain ints."""

    def visit_Subscript(self, node):
        node.value = self.visit(node.value)
        return node

    def visit_Constant(self, node):
        if isinstance(node.value, int) and not isinstance(node.value, bool):
            return ast.copy_location(ast.Call(func=ast.Name(id="_I", ctx=ast.Load()), args=[node], keywords=[]), node)
        return node


def _safe_tree(text):
    """Expression text -> code tree restricted to arithmetic, names, calls of plain names and subscripts.  Attribute access,
    comprehensions, lambdas, strings ... are rejected, so a problem file cannot reach anything but the maths namespace."""
    tree = ast.parse(text.strip(), mode="eval")
    for node in ast.walk(tree):
        if not isinstance(node, _AST_OK):
            raise ValueError(f"unsupported syntax in problem file expression: {type(node).__name__} in {text.strip()[:60]!r}")
        if isinstance(node, ast.Call) and (not isinstance(node.func, ast.Name) or node.keywords):
            raise ValueError(f"only plain function calls are allowed in problem file expressions: {text.strip()[:60]!r}")
        if isinstance(node, ast.Constant) and not isinstance(node.value, (int, float)):
            raise ValueError(f"unsupported literal in problem file expression: {node.value!r}")
        if isinstance(node, ast.Name) and node.id.startswith("__") and not node.id.startswith("__aux_"):
            raise ValueError(f"unsupported name {node.id!r} in problem file expression")
    return ast.fix_missing_locations(_ExactInts().visit(tree))


def load_mac(path, name=None):
    text = open(path).read()
    stmts = _statements(text)
    name = name or re.sub(r"^optDef", "", re.sub(r"\.mac$", "", path.split("/")[-1]))
    P = Problem(name)

    # sign assumptions first: they decide how symbols are created
    assume = {}
    for s in stmts:
        m = re.fullmatch(r"assume\(\s*(\w+)\s*([<>])\s*0\s*\)", s)
        if m:
            assume[m.group(1)] = {"positive": True} if m.group(2) == ">" else {"negative": True}

    ns = dict(_FUNCS)
    ns["k"] = sp.Symbol("k", integer=True)
    values, helpers, arrays, scalars = {}, {}, {}, {}

    def sym_for(nm):
        if nm in ns:
            return ns[nm]
        if nm not in scalars:
            scalars[nm] = sp.Symbol(nm, real=True, **assume.get(nm, {}))
        return scalars[nm]

    def parse(expr, local=None):
        """Maxima expression text -> sympy, creating parameters on demand."""
        e = expr.replace("^", "**")
        e = re.sub(r"'(\w+)", r"__aux_\1", e)
        env = dict(ns)
        env.update(local or {})
        for nm in set(re.findall(r"\b([A-Za-z_]\w*)\s*\[", e)):
            if nm not in env:
                env[nm] = arrays.setdefault(nm, _ParamArray(nm, assume.get(nm, {})))
        for nm in set(re.findall(r"\b([A-Za-z_]\w*)\b", e)):
            if nm in env or nm.startswith("__aux_"):
                continue
            if nm in values:                 # plain reference to a value: substitute it (Maxima evaluates)
                env[nm] = values[nm][0]
            elif nm in helpers:
                env[nm] = helpers[nm]
            elif not re.fullmatch(r"\d+(\.\d*)?([eE][+-]?\d+)?", nm) and nm not in ("e", "E"):
                env[nm] = sym_for(nm)
        for nm in set(re.findall(r"__aux_(\w+)", e)):
            if nm not in values:
                raise ValueError(f"'{nm} used before it is defined")
            env["__aux_" + nm] = values[nm][1]
        env["integrate"] = lambda ex, var, a, b: sp.integrate(ex, (var, a, b))
        env["_I"] = sp.Integer
        return sp.sympify(eval(compile(_safe_tree(e), "<mac>", "eval"), {"__builtins__": {}}, env))  # noqa: S307 (whitelisted AST)

    lists = {}
    for s in stmts:
        if s.startswith("assume(") or s.startswith("load(") or s.startswith("print("):
            continue
        m = re.fullmatch(r"(\w+)\s*\(([^)]*)\)\s*:=\s*(.+)", s, flags=re.S)
        if m:  # helper function
            fname, args, body = m.group(1), [a.strip() for a in m.group(2).split(",")], m.group(3)

            def helper(*vals, _args=args, _body=body):
                return parse(_body, dict(zip(_args, vals)))
            helpers[fname] = helper
            continue
        m = re.fullmatch(r"(\w+)\s*\[\s*(\w+)\s*\]\s*:\s*(.+)", s, flags=re.S)
        if m and m.group(1) in _SPECIAL_ARRAYS:
            arr, idx, body = m.groups()
            lists.setdefault(arr, []).append((idx, parse(body)))
            continue
        m = re.fullmatch(r"(\w+)\s*:\s*(.+)", s, flags=re.S)
        if not m:
            raise ValueError(f"cannot parse statement: {s[:60]}")
        lhs, body = m.groups()
        if lhs in ("x", "u"):
            names = [t.strip() for t in body.strip().strip("[]").split(",")]
            syms = []
            for nm in names:
                sy = sp.Symbol(nm, real=True, **assume.get(nm, {}))
                ns[nm] = sy
                syms.append(sy)
            (P.x if lhs == "x" else P.u).extend(syms)
            continue
        val = parse(body)
        if lhs == "L":
            P.L = val
        elif lhs == "F":
            P.F = val
        else:
            # any other defined symbol is a potential auxiliary value (README.md:38)
            handle = P.def_aux(lhs, val)
            values[lhs] = (val, handle)

    # values that were never referenced with an apostrophe are not auxiliaries after all
    used = set()
    for e in [P.L, P.F] + [v for lst in lists.values() for _, v in lst] + [a.definition for a in P.aux]:
        used |= {f.func.__name__ for f in sp.sympify(e).atoms(sp.Function) if not isinstance(f, tuple(t for t in (sp.sin, sp.cos, sp.tan, sp.asin, sp.acos, sp.Abs)))}
        used |= {s.name.replace("auxsym_", "") for s in sp.sympify(e).free_symbols if s.name.startswith("auxsym_")}
    P.aux = [a for a in P.aux if a.name in used]

    st = {s.name: s for s in P.x}
    for idx, e in lists.get("f", []):
        if idx not in st:
            raise ValueError("elements of f must be indexed by elements of x")
        P.f[st[idx]] = e
    for arr in ("h", "hfe", "hfi", "hle", "hli", "g"):
        items = sorted(lists.get(arr, []), key=lambda t: int(t[0]))   # enforced in ascending index (README.md:36)
        setattr(P, arr, [e for _, e in items])

    # parameters: scalars and arrays, as found in the expressions
    allsyms = set()
    for e in [P.L, P.F] + list(P.f.values()) + P.h + P.hfe + P.hfi + P.hle + P.hli + P.g + [a.definition for a in P.aux]:
        allsyms |= sp.sympify(e).free_symbols
    from .problem import ParamDesc
    for nm, sy in scalars.items():
        if sy in allsyms:
            d = ParamDesc(nm, 1)
            d.symbols = [sy]
            P.params[nm] = d
    for nm, arr in arrays.items():
        if arr.k_sym is not None:
            d = ParamDesc(nm, -1)
            d.symbols = [arr.k_sym]
        else:
            size = max(arr.elems) + 1
            d = ParamDesc(nm, size)
            d.symbols = [arr[i] for i in range(size)]
        P.params[nm] = d
    return P
