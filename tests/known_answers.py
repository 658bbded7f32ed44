"""Known answers the reference's own tests hold for the hot path, restated as data / closed forms:

* the six analytic event models of reference python/tests/test_events.py:40-746 (x(t) in closed form, sx(t) by central
  differences of the closed form with eps = 1e-6, exactly as the reference's test computes its expectation);
* the literal tables of reference tests/cpp/unittests/testSplines.cpp:66-720 (HermiteSpline values / sensitivities).

Used by tests/test_cpu.py (host build of the solver source) and tests/test_gpu.py (CUDA engine through the C ABI).
"""
from copy import deepcopy

import numpy as np


def _flow(s, a, beta, gamma, eta, x):
    """exp(s A) x for A = [[a, 0, 0], [beta, -gamma, 0], [0, 0, -eta]] in closed form (the reference's test calls mpmath.expm,
    python/tests/test_events.py:742-748; the matrix is triangular, so the flow is elementary and exact)"""
    e1, e2, e3 = np.exp(a * s), np.exp(-gamma * s), np.exp(-eta * s)
    z = (a + gamma) * s   # (e1 - e2) / (a + gamma) = e2 * s * expm1(z) / z: no cancellation at the resonance a = -gamma (the nominal point)
    phi = e2 * s * (np.expm1(z) / z if z != 0.0 else 1.0)
    return np.array([x[0] * e1, x[1] * e2 + beta * x[0] * phi, x[2] * e3])


# ---------------------------------------------------------------------------------------------------------------------------
# event models: name -> (parameters in model order, timepoints, x_expected(t, **parameters))
# ---------------------------------------------------------------------------------------------------------------------------
def _x_heavisides(t, k1, k2, k3, alpha, beta, gamma, delta, eta, zeta):
    # reference python/tests/test_events.py:178-233
    x0 = np.array([k1, k2, k3], float)
    t1, t2, t3 = (np.log(k3) - np.log(k1)) / eta, delta, zeta

    def early(s):
        if s < t1:
            return _flow(s, 0.0, beta, gamma, eta, x0)
        xa = _flow(t1, 0.0, beta, gamma, eta, x0)
        xa = xa + np.array([-xa[2] / 2, 0, 0])
        return _flow(s - t1, 0.0, beta, gamma, eta, xa)

    if t < t2:
        return early(t)
    x2 = early(t2)
    if t < t3:
        return _flow(t - t2, -alpha, beta, gamma, eta, x2)
    x3 = _flow(t3 - t2, -alpha, beta, gamma, eta, x2) + np.array([0, 0, zeta / 3])
    return _flow(t - t3, -alpha, beta, gamma, eta, x3) + np.array([0, 0, -np.exp(-eta * (t - t3)) / eta + 1 / eta])


def _x_nested(t, k1, k2, inflow_1, decay_1, decay_2, bolus):
    # reference python/tests/test_events.py:316-351
    equil = inflow_1 / decay_1
    te = (-1 / decay_1) * np.log((inflow_1 / decay_2 - inflow_1 / decay_1) / (k1 - inflow_1 / decay_1))

    def early(s):
        return np.array([equil * (1 - np.exp(-decay_1 * s)) + k1 * np.exp(-decay_1 * s), k2 * np.exp(-decay_2 * s)])

    if t < te:
        return early(t)
    xt = early(te)
    a1, a2 = xt[0] + bolus, xt[1] - 1 / te + bolus
    return np.array([equil * (1 - np.exp(decay_1 * (te - t))) + np.exp(decay_1 * te) * a1 * np.exp(-decay_1 * t),
                     a2 * np.exp(decay_2 * te) * np.exp(-decay_2 * t)])


def _x_pw_simple(t, alpha, beta, gamma, x_1_0):
    # reference python/tests/test_events.py:413-424
    if t < alpha:
        return np.array([x_1_0])
    if t < beta:
        return np.array([gamma + t - alpha])
    return np.array([gamma + beta - alpha + 2.5])


def _x_ddeltax(t, alpha, beta, gamma, delta, x_1_0):
    # reference python/tests/test_events.py:486-505
    if t < alpha:
        return np.array([x_1_0 + t])
    if t < beta:
        return np.array([(x_1_0 + alpha) * alpha + (t - alpha)])
    if t < gamma:
        return np.array([((x_1_0 + alpha) * alpha + (beta - alpha)) * delta + (t - beta)])
    return np.array([(((x_1_0 + alpha) * alpha + (beta - alpha)) * delta + (gamma - beta)) ** 2 * 2 + (t - gamma)])


def _x_pw_semi(t, alpha, beta, gamma, x_1_0, x_2_0, delta, eta):
    # reference python/tests/test_events.py:572-594
    t1, t2, hv = alpha / 2, beta, alpha
    x2 = x_2_0 * np.exp(-eta * t)
    if t < t1:
        x1 = x_1_0 * np.exp(-t)
    elif t < hv:
        x1 = gamma * np.exp(-(t - t1))
    elif t < t2:
        x1 = gamma * np.exp(-(hv - t1)) * np.exp(delta * (t - hv))
    else:
        at2 = gamma * np.exp(-(hv - t1)) * np.exp(delta * (t2 - hv)) + x_2_0 * np.exp(-eta * t2)
        x1 = at2 * np.exp(-(t - t2))
    return np.array([x1, x2])


def _x_pw_state(t, alpha, beta, gamma, x_1_0, x_2_0):
    # reference python/tests/test_events.py:661-682
    hv, t1, t2 = alpha, alpha + 1.4 - x_1_0, beta
    assert hv < t1
    x2 = x_2_0 * np.exp(-t)
    if t < hv:
        x1 = x_1_0
    elif t < t1:
        x1 = (t - hv) + x_1_0
    elif t < t2:
        x1 = gamma + (t - hv) + x_1_0
    else:
        x1 = gamma + (t2 - hv) + x_1_0 + x_2_0 * np.exp(-t2)
    return np.array([x1, x2])


EVENT_MODELS = {
    "ev_heavisides": (dict(k1=2, k2=0.01, k3=5, alpha=2, beta=3, gamma=2, delta=3, eta=1, zeta=5), np.linspace(0, 8, 400), _x_heavisides),
    "ev_nested": (dict(k1=0, k2=0, inflow_1=4, decay_1=2, decay_2=5, bolus=0), np.linspace(0, 1, 101), _x_nested),
    "ev_pw_simple": (dict(alpha=2, beta=3, gamma=4.5, x_1_0=1), np.linspace(0.0, 5.0, 100), _x_pw_simple),
    "ev_ddeltax": (dict(alpha=1.5, beta=2.5, gamma=3.5, delta=5.5, x_1_0=1), np.linspace(0.0, 5.0, 100), _x_ddeltax),
    "ev_pw_semi": (dict(alpha=2, beta=3, gamma=4.5, x_1_0=1, x_2_0=5, delta=2.5, eta=1.4), np.linspace(0.0, 5.0, 100), _x_pw_semi),
    "ev_pw_state": (dict(alpha=2, beta=3, gamma=4.5, x_1_0=1, x_2_0=5), np.linspace(0.0, 5.0, 100), _x_pw_state),
}
EVENT_MODEL_DIMS = {"ev_heavisides": dict(nx=3, np=9, nk=0, ny=3), "ev_nested": dict(nx=2, np=6, nk=0, ny=2),
                    "ev_pw_simple": dict(nx=1, np=4, nk=0, ny=1), "ev_ddeltax": dict(nx=1, np=5, nk=0, ny=1),
                    "ev_pw_semi": dict(nx=2, np=7, nk=0, ny=2), "ev_pw_state": dict(nx=2, np=5, nk=0, ny=2)}


def event_model_expectation(name, parameters=None):
    """(ts, x[nt][nx], sx[nt][np][nx]) as the reference's test computes them (test_events.py:709-718)"""
    par0, ts, fx = EVENT_MODELS[name]
    par = dict(par0 if parameters is None else parameters)
    x = np.array([fx(t, **par) for t in ts])
    eps = 1e-6
    sx = []
    for t in ts:
        row = []
        for ip in par:
            pp = deepcopy(par)
            pp[ip] += eps
            up = np.array(fx(t, **pp))
            pp[ip] -= 2 * eps
            row.append((up - np.array(fx(t, **pp))) / (2 * eps))
        sx.append(row)
    return ts, x, np.array(sx)


# tolerances of the reference's checks (python/tests/util.py:46-85)
EVENT_TOL = dict(nosens=dict(solver=(1e-12, 1e-15), x=(5e-9, 1e-13)),
                 fsa=dict(solver=(1e-13, 1e-15), x=(1e-10, 1e-12), sx=(1e-7, 1e-9)))
