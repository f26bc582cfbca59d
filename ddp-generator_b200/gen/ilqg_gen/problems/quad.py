"""Synthetic quadrotor (BASELINE config 5; nothing like it exists in the reference): 12 states
(position, velocity, roll/pitch/yaw, body rates), 4 rotor thrusts, explicit-Euler discretisation, quadratic running
and final costs around a goal state, two box constraints per input.  Trigonometric terms are auxiliary values so that
dynamics and all derivatives share them (README.md:38 of the reference)."""
import sympy as sp

from ..problem import Problem


def define():
    P = Problem("Quad")
    px, py, pz, vx, vy, vz, phi, th, psi, wp, wq, wr = P.states("px py pz vx vy vz phi th psi wp wq wr")
    u = P.inputs("u1 u2 u3 u4")
    dt = P.param("dt")
    mass = P.param("mass")
    grav = P.param("grav")
    J = P.param_array("J", 3)
    arm = P.param("arm")
    kq = P.param("kq")
    cx = P.param_array("cx", 12)
    cf = P.param_array("cf", 12)
    cu = P.param_array("cu", 4)
    xg = P.param_array("xg", 12)
    uh = P.param("uh")
    ulim = P.param_array("ulim", 2)

    sph = P.def_aux("sph", sp.sin(phi)); cph = P.def_aux("cph", sp.cos(phi))
    sth = P.def_aux("sth", sp.sin(th)); cth = P.def_aux("cth", sp.cos(th))
    sps = P.def_aux("sps", sp.sin(psi)); cps = P.def_aux("cps", sp.cos(psi))
    acc = P.def_aux("acc", (u[0] + u[1] + u[2] + u[3]) / mass)          # thrust acceleration along body z

    P.f[px] = px + dt * vx
    P.f[py] = py + dt * vy
    P.f[pz] = pz + dt * vz
    P.f[vx] = vx + dt * acc * (cph * sth * cps + sph * sps)
    P.f[vy] = vy + dt * acc * (cph * sth * sps - sph * cps)
    P.f[vz] = vz + dt * (acc * cph * cth - grav)
    P.f[phi] = phi + dt * (wp + (sph * wq + cph * wr) * sth / cth)
    P.f[th] = th + dt * (cph * wq - sph * wr)
    P.f[psi] = psi + dt * (sph * wq + cph * wr) / cth
    P.f[wp] = wp + dt * (arm * (u[1] - u[3]) - (J[2] - J[1]) * wq * wr) / J[0]
    P.f[wq] = wq + dt * (arm * (u[2] - u[0]) - (J[0] - J[2]) * wp * wr) / J[1]
    P.f[wr] = wr + dt * (kq * (u[0] - u[1] + u[2] - u[3]) - (J[1] - J[0]) * wp * wq) / J[2]

    xs = [px, py, pz, vx, vy, vz, phi, th, psi, wp, wq, wr]
    P.L = sum(cx[i] * (xs[i] - xg[i]) ** 2 for i in range(12)) + sum(cu[j] * (u[j] - uh) ** 2 for j in range(4))
    P.F = sum(cf[i] * (xs[i] - xg[i]) ** 2 for i in range(12))
    P.h = []
    for j in range(4):
        P.h += [-u[j] + ulim[0], u[j] - ulim[1]]
    return P
