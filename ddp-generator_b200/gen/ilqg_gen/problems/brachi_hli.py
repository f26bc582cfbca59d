"""Brachistochrone with a running state inequality and a [k]-indexed parameter -- restates
examples/Brachistochrone/optDefBrachi_hli.mac:1-14 of the reference (hli[1]: ymin[k]-y, hfe[1]: y-ymin[k])."""
import sympy as sp

from ..problem import Problem


def define():
    P = Problem("BrachiHli")
    (y,) = P.states("y")
    (dy,) = P.inputs("dy")
    dx = P.param("dx")
    g = P.param("g")
    ymin = P.param_k("ymin")
    P.f[y] = y + dy * dx
    P.L = sp.sqrt(2 * (1 + dy**2) / g) * (sp.sqrt(-y) - sp.sqrt(-y - dx * dy)) / dy
    P.F = sp.Integer(0)
    P.hli = [ymin - y]
    P.hfe = [y - ymin]
    return P
