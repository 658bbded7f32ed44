"""Symbolic definitions of the BASELINE models, fed to the CUDA emitter.

They restate the model definitions of the reference's test-model module
(reference python/sdist/amici/testing/models.py: JAK-STAT :338-443,
steadystate :483-525) in SymPy; the pre-generated C++ of the same models
(reference models/model_*_py/) is what the compiled oracle runs, so every
expression here is checked numerically against it by the parity tests.
"""
from __future__ import annotations

import os

import sympy as sp

from .emitter import ModelSpec, Observable, Spline, write_model


def _syms(names):
    return [sp.Symbol(n, real=True) for n in names.split()]


def jakstat_adjoint() -> ModelSpec:
    STAT, pSTAT, pSTAT_pSTAT, npSTAT_npSTAT, nSTAT1, nSTAT2, nSTAT3, nSTAT4, nSTAT5 = states = _syms(
        "STAT pSTAT pSTAT_pSTAT npSTAT_npSTAT nSTAT1 nSTAT2 nSTAT3 nSTAT4 nSTAT5")
    params = _syms("p1 p2 p3 p4 init_STAT sp1 sp2 sp3 sp4 sp5 offset_tSTAT offset_pSTAT scale_tSTAT "
                   "scale_pSTAT sigma_pSTAT sigma_tSTAT sigma_pEpoR")
    (p1, p2, p3, p4, init_STAT, sp1, sp2, sp3, sp4, sp5, offset_tSTAT, offset_pSTAT, scale_tSTAT, scale_pSTAT,
     sigma_pSTAT, sigma_tSTAT, sigma_pEpoR) = params
    Omega_cyt, Omega_nuc = fixed = _syms("Omega_cyt Omega_nuc")
    u = sp.Symbol("u", real=True)
    xdot = [
        (Omega_nuc * p4 * nSTAT5 - Omega_cyt * STAT * p1 * u) / Omega_cyt,
        STAT * p1 * u - 2 * p2 * pSTAT ** 2,
        p2 * pSTAT ** 2 - p3 * pSTAT_pSTAT,
        -(Omega_nuc * p4 * npSTAT_npSTAT - Omega_cyt * p3 * pSTAT_pSTAT) / Omega_nuc,
        -p4 * (nSTAT1 - 2 * npSTAT_npSTAT),
        p4 * (nSTAT1 - nSTAT2),
        p4 * (nSTAT2 - nSTAT3),
        p4 * (nSTAT3 - nSTAT4),
        p4 * (nSTAT4 - nSTAT5),
    ]
    x0 = [init_STAT, 0, 0, 0, 0, 0, 0, 0, 0]
    obs = [
        Observable("obs_pSTAT", offset_pSTAT + scale_pSTAT / init_STAT * (pSTAT + 2 * pSTAT_pSTAT), sigma_pSTAT),
        Observable("obs_tSTAT", offset_tSTAT + scale_tSTAT / init_STAT * (STAT + pSTAT + 2 * pSTAT_pSTAT),
                   sigma_tSTAT),
        Observable("obs_spline", u, sigma_pEpoR),
    ]
    p_nom = [10 ** 0.60, 10 ** 3, 10 ** -0.95, 10 ** -0.0075, 10 ** 0, 10 ** -2.8, 10 ** -0.26, 10 ** -0.075,
             10 ** -0.41, 10 ** -5, 10 ** -0.74, 10 ** -0.64, 10 ** -0.11, 10 ** 0.027, 10 ** -0.5, 10 ** 0,
             10 ** -0.5]
    return ModelSpec(
        name="jakstat_adjoint", states=states, params=params, fixed=fixed, xdot=xdot, x0=x0, observables=obs,
        p_nominal=p_nom, k_nominal=[1.4, 0.45],
        splines=[Spline("u", [0, 5, 10, 20, 60], [sp1, sp2, sp3, sp4, sp5], logarithmic=True)],
    )


def steadystate() -> ModelSpec:
    x1, x2, x3 = states = _syms("x1 x2 x3")
    p1, p2, p3, p4, p5 = params = _syms("p1 p2 p3 p4 p5")
    k1, k2, k3, k4 = fixed = _syms("k1 k2 k3 k4")
    xdot = [
        -2 * p1 * x1 ** 2 - p2 * x1 * x2 + 2 * p3 * x2 + p4 * x3 + p5,
        p1 * x1 ** 2 - p2 * x1 * x2 - p3 * x2 + p4 * x3,
        p2 * x1 * x2 - p4 * x3 - k4 * x3,
    ]
    obs = [Observable("obs_x1", x1), Observable("obs_x2", x2), Observable("obs_x3", x3)]
    return ModelSpec(name="steadystate", states=states, params=params, fixed=fixed, xdot=xdot, x0=[k1, k2, k3],
                     observables=obs, p_nominal=[1, 0.5, 0.4, 2, 0.1], k_nominal=[0.1, 0.4, 0.7, 1])


def robertson_ode() -> ModelSpec:
    """Robertson's problem in ODE form (the shipped reference model is the index-1 DAE run by
    IDAS, reference models/model_robertson_py/; BASELINE config 1 asks for CVODES BDF)."""
    x1, x2, x3 = states = _syms("x1 x2 x3")
    p1, p2, p3 = params = _syms("p1 p2 p3")
    (k1,) = fixed = _syms("k1")
    xdot = [-p1 * x1 + p2 * x2 * x3, p1 * x1 - p2 * x2 * x3 - p3 * x2 ** 2, p3 * x2 ** 2]
    obs = [Observable("obs_x1", x1), Observable("obs_x2", 1e4 * x2), Observable("obs_x3", x3)]
    return ModelSpec(name="robertson_ode", states=states, params=params, fixed=fixed, xdot=xdot,
                     x0=[k1, 0, 1 - k1], observables=obs, p_nominal=[0.04, 1e4, 3e7], k_nominal=[0.9])


ALL = {"jakstat_adjoint": jakstat_adjoint, "steadystate": steadystate, "robertson_ode": robertson_ode}


def generate_all(out_dir: str | None = None) -> list[str]:
    out_dir = out_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "models")
    return [write_model(f(), out_dir) for f in ALL.values()]


if __name__ == "__main__":
    for p in generate_all():
        print("wrote", p)
