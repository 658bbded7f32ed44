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

from .emitter import Event, EventObservable, ModelSpec, Observable, Spline, t_sym, write_model


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


def events() -> ModelSpec:
    """The reference's events test model (reference python/sdist/amici/testing/models.py:261-323; generated code
    models/model_events_py/): two state-triggered events with event observables z = t, and two piecewise time
    switches that the reference importer turns into +/- root pairs (root.cpp: p4-t, t-p4, 4-t, t-4) whose
    Heaviside states h[2], h[4] enter xdot."""
    x1, x2, x3 = states = _syms("x1 x2 x3")
    p1, p2, p3, p4 = params = _syms("p1 p2 p3 p4")
    k1, k2, k3, k4 = fixed = _syms("k1 k2 k3 k4")
    h2, h4 = _syms("Heaviside_2 Heaviside_4")
    t = t_sym
    xdot = [-p1 * x1 * (1 - h2), p2 * x1 * sp.exp(-t / 10) - p3 * x2, -h4 - x3 + 1]
    evs = [Event(x2 - x3), Event(x1 - x3), Event(p4 - t), Event(t - p4), Event(4 - t), Event(t - 4)]
    zs = [EventObservable("z1", 0, t), EventObservable("z2", 1, t)]
    return ModelSpec(name="events", states=states, params=params, fixed=fixed, xdot=xdot, x0=[k1, k2, k3],
                     observables=[Observable("y1", p4 * (x1 + x2 + x3))], p_nominal=[0.5, 2, 0.5, 0.5],
                     k_nominal=[4, 8, 10, 4], events=evs, event_observables=zs, heavisides={h2: 2, h4: 4})


def neuron() -> ModelSpec:
    """The reference's neuron test model (reference python/sdist/amici/testing/models.py:192-258; generated code
    models/model_neuron_py/): Izhikevich neuron, spike event v > 30 with reset v <- -c, u <- u + d and z = t."""
    v, u = states = _syms("v u")
    a, b, c, d = params = _syms("a b c d")
    v0, I0 = fixed = _syms("v0 I0")
    xdot = [sp.Rational(1, 25) * v ** 2 + 5 * v + 140 - u + I0, a * (b * v - u)]
    evs = [Event(v - 30, assignments={v: -c, u: d + u})]
    return ModelSpec(name="neuron", states=states, params=params, fixed=fixed, xdot=xdot, x0=[v0, b * v0],
                     observables=[Observable("y1", v)], p_nominal=[0.02, 0.3, 65, 0.9], k_nominal=[-60, 10],
                     events=evs, event_observables=[EventObservable("z1", 0, t_sym)])


# ---- the reference's closed-form event models (reference python/tests/test_events.py:40-746) --------------------------------------
# Known-answer models: x(t) and (by central differences of the closed form) sx(t) are analytic.  The reference builds them from
# SBML rate rules / events; here the same ODEs, triggers and assignments in SymPy.  All SBML parameters are free parameters
# (initial values included), there are no fixed parameters.  A `piecewise(a, cond, b)` of the SBML model is written with the
# Heaviside state of the event whose root is the condition (the reference importer does the same, cf. models/model_events_py).
def _ev_model(name, states, params, nominal, xdot, x0, evs, heavisides=None) -> ModelSpec:
    obs = [Observable(f"obs_{s}", s) for s in states]
    return ModelSpec(name=name, states=states, params=params, fixed=[], xdot=xdot, x0=x0, observables=obs, p_nominal=nominal,
                     k_nominal=[], events=evs, heavisides=heavisides)


def ev_heavisides() -> ModelSpec:
    """events_plus_heavisides (reference python/tests/test_events.py:116-262)"""
    x1, x2, x3 = states = _syms("x_1 x_2 x_3")
    k1, k2, k3, alpha, beta, gamma, delta, eta, zeta = params = _syms("k1 k2 k3 alpha beta gamma delta eta zeta")
    h_delta, h_zeta = _syms("H_delta H_zeta")
    t = t_sym
    xdot = [-alpha * x1 * h_delta, beta * x1 - gamma * x2, -eta * x3 + h_zeta]
    evs = [Event(k1 - x3, assignments={x1: x1 - x3 / 2}, initial_value=False), Event(t - zeta, assignments={x3: x3 + zeta / 3}, initial_value=False),
           Event(t - delta, initial_value=False)]
    return _ev_model("ev_heavisides", states, params, [2, 0.01, 5, 2, 3, 2, 3, 1, 5], xdot, [k1, k2, k3], evs, {h_delta: 2, h_zeta: 1})


def ev_nested() -> ModelSpec:
    """nested_events (reference python/tests/test_events.py:265-377): the second event is triggered by the first one's bolus"""
    x1, x2 = states = _syms("x_1 x_2")
    k1, k2, inflow_1, decay_1, decay_2, bolus = params = _syms("k1 k2 inflow_1 decay_1 decay_2 bolus")
    t = t_sym
    xdot = [inflow_1 - decay_1 * x1, -decay_2 * x2]
    evs = [Event(x1 - inflow_1 / decay_2, assignments={x2: x2 - 1 / t}, initial_value=False),
           Event(-sp.Rational(1, 2) - x2, assignments={x1: x1 + bolus, x2: x2 + bolus}, initial_value=False)]
    return _ev_model("ev_nested", states, params, [0, 0, 4, 2, 5, 0], xdot, [k1, k2], evs)


def ev_pw_simple() -> ModelSpec:
    """piecewise_plus_event_simple_case (reference python/tests/test_events.py:380-450)"""
    (x1,) = states = _syms("x_1")
    alpha, beta, gamma, x_1_0 = params = _syms("alpha beta gamma x_1_0")
    h_a, h_b = _syms("H_alpha H_beta")
    t = t_sym
    evs = [Event(t - alpha, assignments={x1: gamma}, initial_value=False), Event(t - beta, assignments={x1: x1 + sp.Float(2.5)}, initial_value=False)]
    return _ev_model("ev_pw_simple", states, params, [2, 3, 4.5, 1], [h_a * (1 - h_b)], [x_1_0], evs, {h_a: 0, h_b: 1})


def ev_ddeltax() -> ModelSpec:
    """event_state_dep_ddeltax_dtpx (reference python/tests/test_events.py:453-535): boluses whose derivatives w.r.t. time,
    parameters and states depend on the state"""
    (x1,) = states = _syms("x_1")
    alpha, beta, gamma, delta, x_1_0 = params = _syms("alpha beta gamma delta x_1_0")
    t = t_sym
    evs = [Event(t - alpha, assignments={x1: x1 * t}, initial_value=False), Event(t - beta, assignments={x1: x1 * delta}, initial_value=False),
           Event(t - gamma, assignments={x1: 2 * x1 * x1}, initial_value=False)]
    return _ev_model("ev_ddeltax", states, params, [1.5, 2.5, 3.5, 5.5, 1], [sp.Integer(1)], [x_1_0], evs)


def ev_pw_semi() -> ModelSpec:
    """piecewise_plus_event_semi_complicated (reference python/tests/test_events.py:538-620)"""
    x1, x2 = states = _syms("x_1 x_2")
    alpha, beta, gamma, x_1_0, x_2_0, delta, eta = params = _syms("alpha beta gamma x_1_0 x_2_0 delta eta")
    h_a, h_b = _syms("H_alpha H_beta")
    t = t_sym
    inside = h_a * (1 - h_b)
    xdot = [delta * x1 * inside - x1 * (1 - inside), -eta * x2]
    evs = [Event(t - alpha / 2, assignments={x1: gamma}, initial_value=False), Event(t - beta, assignments={x1: x1 + x2}, initial_value=False),
           Event(t - alpha, initial_value=False)]
    return _ev_model("ev_pw_semi", states, params, [2, 3, 4.5, 1, 5, 2.5, 1.4], xdot, [x_1_0, x_2_0], evs, {h_a: 2, h_b: 1})


def ev_pw_state() -> ModelSpec:
    """piecewise_plus_event_trigger_depends_on_state (reference python/tests/test_events.py:623-706)"""
    x1, x2 = states = _syms("x_1 x_2")
    alpha, beta, gamma, x_1_0, x_2_0 = params = _syms("alpha beta gamma x_1_0 x_2_0")
    h_a, h_b = _syms("H_alpha H_beta")
    t = t_sym
    evs = [Event(x1 - sp.Float(1.4), assignments={x1: x1 + gamma}, initial_value=False), Event(t - beta, assignments={x1: x1 + x2}, initial_value=False),
           Event(t - alpha, initial_value=False)]
    return _ev_model("ev_pw_state", states, params, [2, 3, 4.5, 1, 5], [h_a * (1 - h_b), -x2], [x_1_0, x_2_0], evs, {h_a: 2, h_b: 1})


def spline_modes() -> ModelSpec:
    """Two forcing splines outside the default HermiteSpline configuration (csrc/hermite_spline.cuh, spline.cuh general path):
    u1 — equidistant nodes, finite-difference slopes, `given` boundaries, linear extrapolation on both sides, logarithmic;
    u2 — explicit nodes, slopes given as parameter expressions, periodic boundaries and periodic extrapolation, not logarithmic
    (the reference's periodic extrapolation of a LOGARITHMIC spline exponentiates twice outside the node range — reproduced by
    hermite_spline.cuh and covered by the random spline cases — which would put a jump into the forcing here).
    The output grid starts before the first node and ends after the last one."""
    x1, x2 = states = _syms("x1 x2")
    kd, a0, a1, a2, a3, b0, b1, b2, d = params = _syms("kd a0 a1 a2 a3 b0 b1 b2 d")
    (x10,) = fixed = _syms("x10")
    u1, u2 = _syms("u1 u2")
    xdot = [-kd * x1 + u1, -x2 + u2]
    obs = [Observable("obs_x1", x1), Observable("obs_x2", x2), Observable("obs_u2", u2)]
    return ModelSpec(
        name="spline_modes", states=states, params=params, fixed=fixed, xdot=xdot, x0=[x10, sp.Integer(0)], observables=obs,
        p_nominal=[0.7, 1.0, 2.5, 0.6, 1.4, 0.8, 2.0, 0.5, 0.3], k_nominal=[1.0],
        splines=[Spline("u1", [1.0, 4.0], [a0, a1, a2, a3], logarithmic=True, bc=("given", "given"), extrapolate=("linear", "linear"),
                        equidistant=True),
                 Spline("u2", [0.5, 1.25, 2.75, 3.5], [b0, b1, b2, b0], logarithmic=False, derivatives=[d, -d, d / 2, d],
                        bc=("periodic", "periodic"), extrapolate=("periodic", "periodic"))])


ANALYTIC = {"ev_heavisides": ev_heavisides, "ev_nested": ev_nested, "ev_pw_simple": ev_pw_simple, "ev_ddeltax": ev_ddeltax,
            "ev_pw_semi": ev_pw_semi, "ev_pw_state": ev_pw_state}


ALL = {"jakstat_adjoint": jakstat_adjoint, "steadystate": steadystate, "robertson_ode": robertson_ode,
       "events": events, "neuron": neuron, **ANALYTIC, "spline_modes": spline_modes}


def generate_all(out_dir: str | None = None) -> list[str]:
    out_dir = out_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "models")
    return [write_model(f(), out_dir) for f in ALL.values()]


if __name__ == "__main__":
    for p in generate_all():
        print("wrote", p)
