"""Brachistochrone -- restates examples/Brachistochrone/optDefBrachi.mac:1-13 of the reference.

The running cost is the travel time along one straight segment, i.e. the closed form of
integrate(sqrt((1+dy^2)/(2*g*abs(y+x_*dy))), x_, 0, dx) under the file's assumptions dx>0, y<0, dy<0.
"""
import sympy as sp

from ..problem import Problem


def define():
    P = Problem("Brachi")
    (y,) = P.states("y")
    (dy,) = P.inputs("dy")
    dx = P.param("dx")
    g = P.param("g")
    yf = P.param("yf")
    P.f[y] = y + dy * dx
    P.L = sp.sqrt(2 * (1 + dy**2) / g) * (sp.sqrt(-y) - sp.sqrt(-y - dx * dy)) / dy
    P.F = sp.Integer(0)
    P.hfe = [y - yf]
    return P
