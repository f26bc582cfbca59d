"""Twin-motor pendulum with a running EQUALITY constraint and a terminal INEQUALITY constraint -- the two constraint
kinds of the reference's augmented-Lagrangian family that none of its shipped examples uses (hle: folding
genenerator_main.mac:93-107, update iLQG_func.tem:427-453; hfi: folding genenerator_main.mac:58-72, update
iLQG_func.tem:470-509).  Authored here (nothing in the reference): two motors drive one joint and must share the
load equally (hle[1]: ua - ub = 0), the final angle must stay above thmin (hfi[1]: thmin - th <= 0) while the costs
pull it to zero, and both motors are torque limited (h[1..4])."""
import sympy as sp

from ..problem import Problem


def define():
    P = Problem("Pend")
    th, om = P.states("th om")
    ua, ub = P.inputs("ua ub")
    dt = P.param("dt")
    gl = P.param("gl")
    cu = P.param_array("cu", 2)
    cx = P.param_array("cx", 2)
    cf = P.param_array("cf", 2)
    lim = P.param_array("lim", 2)
    thmin = P.param("thmin")
    P.f[th] = th + dt * om
    P.f[om] = om + dt * (-gl * sp.sin(th) + ua + ub)
    P.L = cu[0] * ua**2 + cu[1] * ub**2 + cx[0] * th**2 + cx[1] * om**2
    P.F = cf[0] * th**2 + cf[1] * om**2
    P.h = [-ua + lim[0], ua - lim[1], -ub + lim[0], ub - lim[1]]
    P.hle = [ua - ub]
    P.hfi = [thmin - th]
    # user outputs for calcG (iLQG_func.tem:511-521; no reference example defines any): pendulum energy and net torque
    P.g = [om**2 / 2 - gl * sp.cos(th), ua + ub]
    return P
