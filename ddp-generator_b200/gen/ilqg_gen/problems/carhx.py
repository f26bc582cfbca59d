"""Car parking with a STATE-DEPENDENT steering limit: |w| <= limW / (1 + kv v^2).  Exercises the reference's
extension "state dependent input constraints" (README.md:13; genenerator_main.mac:373-447; back_pass.c:183-199):
the active constraint's gradient hx enters the feedback gains of clamped inputs."""
import sympy as sp

from ..problem import Problem
from .car import sqrt_abs


def define():
    P = Problem("CarHx")
    x_, y_, t, v = P.states("x_ y_ t v")
    w, a = P.inputs("w a")
    d = P.param("d")
    h = P.param("h")
    kv = P.param("kv")
    cf = P.param_array("cf", 4)
    pf = P.param_array("pf", 4)
    cx = P.param_array("cx", 2)
    px = P.param_array("px", 2)
    cu = P.param_array("cu", 2)
    limW = P.param_array("limW", 2)
    limA = P.param_array("limA", 2)
    s = P.def_aux("s", d + h * v * sp.cos(w) - sp.sqrt(d**2 - (h * v * sp.sin(w)) ** 2))
    shrink = P.def_aux("shrink", 1 / (1 + kv * v**2))
    P.f[x_] = x_ + s * sp.cos(t)
    P.f[y_] = y_ + s * sp.sin(t)
    P.f[t] = t + sp.asin(sp.sin(w) * h * v / d)
    P.f[v] = v + h * a
    P.F = (cf[0] * sqrt_abs(x_, pf[0]) + cf[1] * sqrt_abs(y_, pf[1]) + cf[2] * sqrt_abs(t, pf[2])
           + cf[3] * sqrt_abs(v, pf[3]) + cx[0] * sqrt_abs(x_, px[0]) + cx[1] * sqrt_abs(y_, px[1]))
    P.L = cu[0] * w**2 + cu[1] * a**2 + cx[0] * sqrt_abs(x_, px[0]) + cx[1] * sqrt_abs(y_, px[1])
    P.h = [-w + limW[0] * shrink, w - limW[1] * shrink, -a + limA[0], a - limA[1]]
    return P
