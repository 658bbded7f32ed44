"""Test infrastructure: the shipped test models as the REFERENCE's symbolic model objects (`DEModel`), built with the reference's
own component classes (reference python/sdist/amici/_symbolic/de_model_components.py) the way its importers do
(python/sdist/amici/importers/pysb/__init__.py:373-420: `DEModel()` + `add_component`).  The model definitions follow the
reference's test-model module (python/sdist/amici/testing/models.py: steadystate :483-525, neuron :192-258, events :261-323,
JAK-STAT :338-443), which builds them from SBML — not possible here (no libsbml).

Used by tests/test_cpu.py::test_de_export_* to drive amici_b200.codegen.de_export with real `DEModel` objects.
"""
import sympy as sp

from tests import ref_python


def _ref():
    ref_python.load()
    from amici._symbolic import de_model_components as C
    from amici._symbolic.de_model import DEModel
    from amici.importers.utils import amici_time_symbol, symbol_with_assumptions
    return DEModel, C, symbol_with_assumptions, amici_time_symbol


def _base(states, x0, xdot, params, p_nom, fixed, k_nom, obs, sigmas=None, expressions=()):
    DEModel, C, S, _ = _ref()
    m = DEModel()
    for s, v in zip(params, p_nom):
        m.add_component(C.FreeParameter(s, str(s), v))
    for s, v in zip(fixed, k_nom):
        m.add_component(C.FixedParameter(s, str(s), v))
    for s, v in expressions:
        m.add_component(C.Expression(s, str(s), v))
    for s, i, d in zip(states, x0, xdot):
        m.add_component(C.DifferentialState(s, str(s), sp.sympify(i), sp.sympify(d)))
    for j, (name, formula) in enumerate(obs):
        y, my, sig = S(name), S("m" + name), S("sigma_" + name)
        m.add_component(C.Observable(y, name, formula, my))
        m.add_component(C.SigmaY(sig, str(sig), sp.sympify(sigmas[j]) if sigmas else sp.Float(1.0)))
        m.add_component(C.LogLikelihoodY(S("J" + name), "J" + name,
                                         0.5 * sp.log(2 * sp.pi * sig ** 2) + 0.5 * ((y - my) / sig) ** 2))
    return m


def steadystate():
    _, _, S, _ = _ref()
    x1, x2, x3 = states = [S(n) for n in ("x1", "x2", "x3")]
    p1, p2, p3, p4, p5 = params = [S(f"p{i}") for i in range(1, 6)]
    k1, k2, k3, k4 = fixed = [S(f"k{i}") for i in range(1, 5)]
    xdot = [-2 * p1 * x1 ** 2 - p2 * x1 * x2 + 2 * p3 * x2 + p4 * x3 + p5,
            p1 * x1 ** 2 - p2 * x1 * x2 - p3 * x2 + p4 * x3,
            p2 * x1 * x2 - p4 * x3 - k4 * x3]
    return _base(states, [k1, k2, k3], xdot, params, [1, 0.5, 0.4, 2, 0.1], fixed, [0.1, 0.4, 0.7, 1],
                 [("obs_x1", x1), ("obs_x2", x2), ("obs_x3", x3)])


def robertson_ode():
    _, _, S, _ = _ref()
    x1, x2, x3 = states = [S(n) for n in ("x1", "x2", "x3")]
    p1, p2, p3 = params = [S(f"p{i}") for i in range(1, 4)]
    (k1,) = fixed = [S("k1")]
    xdot = [-p1 * x1 + p2 * x2 * x3, p1 * x1 - p2 * x2 * x3 - p3 * x2 ** 2, p3 * x2 ** 2]
    return _base(states, [k1, 0, 1 - k1], xdot, params, [0.04, 1e4, 3e7], fixed, [0.9],
                 [("obs_x1", x1), ("obs_x2", 1e4 * x2), ("obs_x3", x3)])


def steadystate_with_expressions():
    """the same model written with intermediate expressions (`w`): flux_r0 ... as the SBML importer produces them, one of
    them depending on another — exercises the inlining of `w`"""
    _, _, S, _ = _ref()
    x1, x2, x3 = states = [S(n) for n in ("x1", "x2", "x3")]
    p1, p2, p3, p4, p5 = params = [S(f"p{i}") for i in range(1, 6)]
    k1, k2, k3, k4 = fixed = [S(f"k{i}") for i in range(1, 5)]
    r0, r1, r2, r3, r23 = [S(n) for n in ("flux_r0", "flux_r1", "flux_r2", "flux_r3", "flux_r23")]
    expressions = [(r23, r2 - r3), (r0, p1 * x1 ** 2), (r1, p2 * x1 * x2), (r2, p3 * x2), (r3, p4 * x3)]
    xdot = [-2 * r0 - r1 + 2 * r2 + r3 + p5, r0 - r1 - r23, r1 - r3 - k4 * x3]
    return _base(states, [k1, k2, k3], xdot, params, [1, 0.5, 0.4, 2, 0.1], fixed, [0.1, 0.4, 0.7, 1],
                 [("obs_x1", x1), ("obs_x2", x2), ("obs_x3", x3)], expressions=expressions)


def neuron():
    _, C, S, t = _ref()
    v, u = states = [S("v"), S("u")]
    a, b, c, d = params = [S(n) for n in "abcd"]
    v0, I0 = fixed = [S("v0"), S("I0")]
    xdot = [sp.Rational(1, 25) * v ** 2 + 5 * v + 140 - u + I0, a * (b * v - u)]
    m = _base(states, [v0, b * v0], xdot, params, [0.02, 0.3, 65, 0.9], fixed, [-60, 10], [("y1", v)])
    ev = S("event_1")
    m.add_component(C.Event(ev, "event_1", v - 30, use_values_from_trigger_time=True, assignments={v: -c, u: d + u}, initial_value=True))
    z, mz, sz = S("z1"), S("mz1"), S("sigma_z1")
    m.add_component(C.EventObservable(z, "z1", t, ev, mz))
    m.add_component(C.SigmaZ(sz, "sigma_z1", sp.Float(1.0)))
    return m


def events():
    """Heaviside switches in the right-hand side: `DEModel.parse_events` turns them into +/- root pairs and replaces them by the
    Heaviside state variables (reference python/sdist/amici/_symbolic/de_model.py:1173-1195, :2434-2475)"""
    _, C, S, t = _ref()
    x1, x2, x3 = states = [S(n) for n in ("x1", "x2", "x3")]
    p1, p2, p3, p4 = params = [S(f"p{i}") for i in range(1, 5)]
    k1, k2, k3, k4 = fixed = [S(f"k{i}") for i in range(1, 5)]
    # piecewise(1, time > p4, 0) as the SBML importer hands it over: 1 - Heaviside(p4 - t) (cf. the generated xdot.cpp / root.cpp
    # of models/model_events_py: roots p4 - t, t - p4, 4 - t, t - 4)
    xdot = [-p1 * x1 * (1 - sp.Heaviside(p4 - t, 1)), p2 * x1 * sp.exp(-t / 10) - p3 * x2, -x3 + (1 - sp.Heaviside(4 - t, 1))]
    m = _base(states, [k1, k2, k3], xdot, params, [0.5, 2, 0.5, 0.5], fixed, [4, 8, 10, 4], [("y1", p4 * (x1 + x2 + x3))])
    e1, e2 = S("event_1"), S("event_2")
    m.add_component(C.Event(e1, "event_1", x2 - x3, use_values_from_trigger_time=True, initial_value=True))
    m.add_component(C.Event(e2, "event_2", x1 - x3, use_values_from_trigger_time=True, initial_value=True))
    for i, ev in enumerate((e1, e2)):
        z, mz, sz = S(f"z{i + 1}"), S(f"mz{i + 1}"), S(f"sigma_z{i + 1}")
        m.add_component(C.EventObservable(z, f"z{i + 1}", t, ev, mz))
        m.add_component(C.SigmaZ(sz, str(sz), sp.Float(1.0)))
    return m


def jakstat_adjoint():
    """with the reference's own `CubicHermiteSpline` object (python/sdist/amici/importers/sbml/splines.py) added through
    `DEModel.add_spline`; the spline's model expression is the opaque function the importer creates (`ode_model_symbol`)"""
    _, _, S, t = _ref()
    from amici.importers.sbml.splines import CubicHermiteSpline
    STAT, pSTAT, pSTAT_pSTAT, npSTAT_npSTAT, nSTAT1, nSTAT2, nSTAT3, nSTAT4, nSTAT5 = states = [S(n) for n in (
        "STAT pSTAT pSTAT_pSTAT npSTAT_npSTAT nSTAT1 nSTAT2 nSTAT3 nSTAT4 nSTAT5").split()]
    params = [S(n) for n in ("p1 p2 p3 p4 init_STAT sp1 sp2 sp3 sp4 sp5 offset_tSTAT offset_pSTAT scale_tSTAT scale_pSTAT "
                             "sigma_pSTAT sigma_tSTAT sigma_pEpoR").split()]
    (p1, p2, p3, p4, init_STAT, sp1, sp2, sp3, sp4, sp5, offset_tSTAT, offset_pSTAT, scale_tSTAT, scale_pSTAT, sigma_pSTAT,
     sigma_tSTAT, sigma_pEpoR) = params
    Omega_cyt, Omega_nuc = fixed = [S("Omega_cyt"), S("Omega_nuc")]
    spline = CubicHermiteSpline(sbml_id="u", evaluate_at=t, nodes=[0, 5, 10, 20, 60], values_at_nodes=[sp1, sp2, sp3, sp4, sp5],
                                extrapolate=("constant", "constant"), logarithmic_parametrization=True)
    u = spline.sbml_id
    xdot = [(Omega_nuc * p4 * nSTAT5 - Omega_cyt * STAT * p1 * u) / Omega_cyt,
            STAT * p1 * u - 2 * p2 * pSTAT ** 2,
            p2 * pSTAT ** 2 - p3 * pSTAT_pSTAT,
            -(Omega_nuc * p4 * npSTAT_npSTAT - Omega_cyt * p3 * pSTAT_pSTAT) / Omega_nuc,
            -p4 * (nSTAT1 - 2 * npSTAT_npSTAT), p4 * (nSTAT1 - nSTAT2), p4 * (nSTAT2 - nSTAT3), p4 * (nSTAT3 - nSTAT4),
            p4 * (nSTAT4 - nSTAT5)]
    p_nom = [10 ** 0.60, 10 ** 3, 10 ** -0.95, 10 ** -0.0075, 10 ** 0, 10 ** -2.8, 10 ** -0.26, 10 ** -0.075, 10 ** -0.41, 10 ** -5,
             10 ** -0.74, 10 ** -0.64, 10 ** -0.11, 10 ** 0.027, 10 ** -0.5, 10 ** 0, 10 ** -0.5]
    obs = [("obs_pSTAT", offset_pSTAT + scale_pSTAT / init_STAT * (pSTAT + 2 * pSTAT_pSTAT)),
           ("obs_tSTAT", offset_tSTAT + scale_tSTAT / init_STAT * (STAT + pSTAT + 2 * pSTAT_pSTAT)), ("obs_spline", u)]
    m = _base(states, [init_STAT, 0, 0, 0, 0, 0, 0, 0, 0], xdot, params, p_nom, fixed, [1.4, 0.45], obs,
              sigmas=[sigma_pSTAT, sigma_tSTAT, sigma_pEpoR])
    # `ode_model_symbol` only asks its importer for the model's free parameters
    from amici.constants import SymbolId
    importer = type("Importer", (), {"_symbols": {SymbolId.FREE_PARAMETER: {p: None for p in params}}})()
    m.add_spline(spline, spline.ode_model_symbol(importer))
    return m


def dirac():
    """the reference's Dirac test model (python/sdist/amici/testing/models.py:21-34, :173-189; shipped sources
    models/model_dirac_py): a bolus `x1 = x1 + 1` at the parameter-dependent time p2.  The engine's header of this model exists
    ONLY through this `DEModel` and amici_b200.codegen.de_export (tests/golden/make_de_models.py)."""
    _, C, S, t = _ref()
    x1, x2 = states = [S("x1"), S("x2")]
    p1, p2, p3, p4 = params = [S(f"p{i}") for i in range(1, 5)]
    m = _base(states, [sp.Float(0.0), sp.Float(0.0)], [-p1 * x1, p3 * x1 - p4 * x2], params, [1, 0.5, 2, 3], [], [], [("obs_x2", x2)])
    m.add_component(C.Event(S("_E0"), "_E0", t - p2, use_values_from_trigger_time=False, assignments={x1: x1 + 1}, initial_value=True))
    return m


def robertson():
    """the shipped index-1 DAE form of Robertson's problem (python/sdist/amici/testing/models.py:90-134; models/model_robertson_py):
    two differential states, the algebraic state x3 with 0 = x1 + x2 + x3 - 1 and the inconsistent start x3 = 0 that
    IDACalcIC repairs"""
    _, C, S, _ = _ref()
    x1, x2, x3 = S("x1"), S("x2"), S("x3")
    p1, p2, p3 = params = [S(f"p{i}") for i in range(1, 4)]
    (k1,) = fixed = [S("k1")]
    m = _base([x1, x2], [k1, sp.Float(0.0)], [-p1 * x1 + p2 * x2 * x3, p1 * x1 - p2 * x2 * x3 - p3 * x2 ** 2], params, [0.04, 1e4, 3e7],
              fixed, [0.9], [("obs_x1", x1), ("obs_x2", 1e4 * x2), ("obs_x3", x3)])
    m.add_component(C.AlgebraicState(x3, "x3", sp.Float(0.0)))
    m.add_component(C.AlgebraicEquation(S("ae_0"), x1 + x2 + x3 - 1))
    return m


def nested_events():
    """the reference's nested-events test model (python/sdist/amici/testing/models.py:446-480; shipped sources
    models/model_nested_events_py): growth switched on by a STATE-dependent Heaviside (Virus > 1) and an injection at the
    parameter-dependent time t_0 that flips the switch — a secondary event of the first"""
    _, C, S, t = _ref()
    (virus,) = states = [S("Virus")]
    V_0, V_0_inject, t_0, rho_V, delta_V = params = [S(n) for n in ("V_0", "V_0_inject", "t_0", "rho_V", "delta_V")]
    xdot = [virus * rho_V * sp.Heaviside(virus - 1, 1) - virus * delta_V]
    m = _base(states, [V_0], xdot, params, [0.1, 1000.0, 2.0, 0.8, 1.6], [], [], [("obs_Virus", virus)])
    m.add_component(C.Event(S("injection"), "injection", t - t_0, use_values_from_trigger_time=False,
                            assignments={virus: virus + V_0_inject}, initial_value=True))
    return m


def calvetti():
    """the reference's Calvetti test model (python/sdist/amici/testing/models.py:37-88, :137-160; shipped sources
    models/model_calvetti_py): an index-1 DAE with three differential and three algebraic states, sixteen intermediate
    expressions, no free parameters, and a forcing s(t) = [10 <= t <= 12] that the importer turns into time-triggered
    Heaviside switches"""
    _, C, S, t = _ref()
    V1, V2, V3, f1, f2, f3 = [S(n) for n in ("V1", "V2", "V3", "f1", "f2", "f3")]
    V1ss, R1ss, V2ss, R2ss, V3ss, R3ss = fixed = [S(n) for n in ("V1ss", "R1ss", "V2ss", "R2ss", "V3ss", "R3ss")]
    p2, p3, L1, L2, L3, C1ss, C2ss, C3ss, R1, R2, R3, f0, s_, rV1, rV2, rV3 = [S(n) for n in (
        "p2 p3 L1 L2 L3 C1ss C2ss C3ss R1 R2 R3 f0 s rate_of_V1 rate_of_V2 rate_of_V3").split()]
    Q = sp.Rational
    third = Q(1, 3)
    flow = (R1 + R2) * f1 + (R2 + R3) * f2 + R3 * f3
    expressions = [
        (p2, 1 - R1ss), (p3, 1 - (R1ss + R2ss)),
        (L1, (R1ss * V1ss ** 2) ** third), (L2, (R2ss * V2ss ** 2) ** third), (L3, (R3ss * V3ss ** 2) ** third),
        (C1ss, V1ss / (1 - Q(1, 2) * R1ss)), (C2ss, V2ss / (p2 - Q(1, 2) * R2ss)), (C3ss, V3ss / (p3 - Q(1, 2) * R3ss)),
        (R1, L1 ** 3 / V1 ** 2), (R2, L2 ** 3 / V2 ** 2), (R3, L3 ** 3 / V3 ** 2),
        (f0, 2 / R1 - flow / R1),
        (s_, sp.Heaviside(t - 10, 1) * sp.Heaviside(12 - t, 1)),
        (rV1, Q(1, 31) * ((Q(129, 100) - V1 / V1ss) / (Q(129, 100) - 1) + s_ - 2 * (V1 / (C1ss * flow)))),
        (rV2, Q(1, 163) * ((Q(151, 100) - V2 / V2ss) / (Q(151, 100) - 1) - 2 * (V2 / (C2ss * ((R2 + R3) * f2 + R3 * f3))))),
        (rV3, Q(1, 122) * ((1000000 - V3 / V3ss) / (1000000 - 1) - 2 * (V3 / (C3ss * (R3 * f3))))),
    ]
    m = _base([V1, V2, V3], [V1ss, V2ss, V3ss], [rV1, rV2, rV3], [], [], fixed, [0.29, 0.74, 0.44, 0.08, 0.27, 0.18],
              [("obs_V1", V1), ("obs_V2", V2), ("obs_V3", V3), ("obs_f0", f0), ("obs_f1", f1), ("obs_f2", f2)], expressions=expressions)
    for a in (f1, f2, f3):
        m.add_component(C.AlgebraicState(a, str(a), sp.Float(1.0)))
    for i, e in enumerate((f0 - rV1 - f1, f1 - rV2 - f2, f2 - rV3 - f3)):
        m.add_component(C.AlgebraicEquation(S(f"ae_{i}"), e))
    return m


def build(name):
    """-> the model as the importers hand it to `DEExporter`: `generate_basic_variables()` parses the events (Heaviside
    functions -> root pairs, reference de_model.py:1156-1171) and creates the symbolic identifiers"""
    m = ALL[name]() if name in ALL else globals()[name]()
    m.generate_basic_variables()
    return m


ALL = {"calvetti": calvetti, "nested_events": nested_events, "robertson": robertson, "dirac": dirac, "jakstat_adjoint": jakstat_adjoint, "steadystate": steadystate, "robertson_ode": robertson_ode, "neuron": neuron, "events": events}
