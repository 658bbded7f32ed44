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
    x = np.asarray(x, np.longdouble)
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
    """(ts, x[nt][nx], sx[nt][np][nx]) as the reference's test computes them (test_events.py:709-718): central differences of
    the closed form with eps = 1e-6 — evaluated in extended precision here.  In double precision the difference quotient carries
    |x| * 2^-53 / eps of round-off (3e-7 relative on ev_pw_semi's sx[x_1][x_2_0], measured against the exact derivative), which is
    the size of the reference's own acceptance tolerance; the engine's error there is 4e-12."""
    par0, ts, fx = EVENT_MODELS[name]
    par = {k_: np.longdouble(v) for k_, v in dict(par0 if parameters is None else parameters).items()}
    ts = [np.longdouble(t) for t in np.asarray(ts, float)]
    x = np.array([fx(t, **par) for t in ts])
    eps = np.longdouble(1e-6)
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
    return np.asarray(ts, float), np.asarray(x, float), np.asarray(sx, float)


# tolerances of the reference's checks (python/tests/util.py:46-85)
EVENT_TOL = dict(nosens=dict(solver=(1e-12, 1e-15), x=(5e-9, 1e-13)),
                 fsa=dict(solver=(1e-13, 1e-15), x=(1e-10, 1e-12), sx=(1e-7, 1e-9)))


# ---------------------------------------------------------------------------------------------------------------------------
# Hermite spline: tests/golden/spline_cases.npz (tests/golden/make_spline_cases.py)
# ---------------------------------------------------------------------------------------------------------------------------
def check_spline_cases(golden_path, run):
    """run(cfg, nodes, values, slopes, dv, dm, ts) -> (value[c][nt], sens[c][nt][4], err[c][nt]) is the implementation under test.
    (1) every literal table of reference tests/cpp/unittests/testSplines.cpp:66-720 with the tolerance that test uses
    (ASSERT_DOUBLE_EQ = 4 ulp), its ASSERT_THROW times and its final values / final sensitivities; the compiled reference's
    own outputs stored next to them must pass the same check (pins the oracle wrapper and the parser).
    (2) ~1300 random splines over every valid boundary condition x extrapolation x FD x equidistant x logarithmic
    combination against the reference class, bit for bit (same operations in the same order), incl. where it throws."""
    g = np.load(golden_path)
    ulp4 = 4 * np.finfo(float).eps

    def close(v, e, rt, at):
        v, e = np.asarray(v, float), np.asarray(e, float)
        if rt or at:
            return bool((np.abs(v - e) <= at + rt * np.abs(e)).all())
        return bool(((v == e) | (np.abs(v - e) <= ulp4 * np.maximum(np.abs(v), np.abs(e)))).all())

    def check_literal(val, sens, err, who):
        n_checked = 0
        for i, name in enumerate(g["lit_names"]):
            rt, at = g["lit_tol"][i]
            for j in range(g["lit_ts"].shape[1]):
                k = g["lit_kind"][i, j]
                where = f"{who}: {name} t={g['lit_ts'][i, j]}"
                if k == 1:
                    assert close(val[i, j], g["lit_expect_v"][i, j], rt, at), (where, val[i, j], g["lit_expect_v"][i, j])
                elif k == 2:
                    e = g["lit_expect_s"][i, j]
                    m = ~np.isnan(e)
                    assert close(sens[i, j][m], e[m], rt, at), (where, sens[i, j], e)
                elif k == 3:
                    assert err[i, j] == 1 and np.isnan(val[i, j]), where
                elif k in (4, 5):
                    e = g["lit_expect_v"][i, j]
                    if k == 4:
                        assert (np.isnan(e) and np.isnan(val[i, j])) or close(val[i, j], e, 0, 0), (where, val[i, j], e)
                    es = g["lit_expect_s"][i, j]
                    m = ~np.isnan(es)
                    assert close(sens[i, j][m], es[m], 0, 0), (where, sens[i, j], es)
                else:
                    continue
                n_checked += 1
        return n_checked

    assert check_literal(g["lit_ref_v"], g["lit_ref_s"], g["lit_ref_thrown"], "compiled reference") > 150
    val, sens, err = run(*(g["lit_" + k] for k in ("cfg", "nodes", "values", "slopes", "dv", "dm", "ts")))
    check_literal(val, sens, err, "candidate")
    np.testing.assert_array_equal(val, g["lit_ref_v"])     # also the rows the tables do not list (NaN == NaN)
    np.testing.assert_array_equal(sens, g["lit_ref_s"])
    val, sens, err = run(*(g["rnd_" + k] for k in ("cfg", "nodes", "values", "slopes", "dv", "dm", "ts")))
    np.testing.assert_array_equal(err, g["rnd_ref_thrown"])
    np.testing.assert_array_equal(val, g["rnd_ref_v"])
    np.testing.assert_array_equal(sens, g["rnd_ref_s"])
    return len(g["rnd_cfg"])


def spline_runner(fn):
    """adapts a C entry point with the signature of ab_selftest_spline (include/amici_b200.h) to check_spline_cases"""
    import ctypes as C
    fn.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4
    fn.restype = C.c_int

    def run(cfg, nodes, values, slopes, dv, dm, ts):
        a = [np.ascontiguousarray(x) for x in (cfg, nodes, values, slopes, dv, dm, ts)]
        n, nt = a[6].shape
        val, sens, err = np.zeros((n, nt)), np.zeros((n, nt, 4)), np.zeros((n, nt), np.int32)
        rc = fn(n, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, 4, a[4].ctypes.data, a[5].ctypes.data, nt,
                a[6].ctypes.data, val.ctypes.data, sens.ctypes.data, err.ctypes.data)
        assert rc == 0, rc
        return val, sens, err
    return run


# ---------------------------------------------------------------------------------------------------------------------------
# model `spline_modes` (amici_b200/codegen/models.py): two forcing splines outside the default configuration
# ---------------------------------------------------------------------------------------------------------------------------
SPLINE_MODES_TS = np.linspace(0.0, 6.0, 25)      # starts before both splines' first node, ends after their last


def spline_modes_expectation(spline_run, p, x10):
    """x(t), sx(t) of  x1' = -kd x1 + u1(t), x2' = -x2 + u2(t)  by scipy (LSODA-free: explicit RK at 1e-13) with u1 / u2 and their
    parameter sensitivities evaluated by `spline_run` (an implementation pinned against the reference class by
    check_spline_cases).  p = (kd, a0..a3, b0, b1, b2, d)."""
    from scipy.integrate import solve_ivp
    kd, a, b, d = p[0], p[1:5], p[5:8], p[8]
    cfg = np.array([[4, -1, -1, 1, 1, 1, 1, 1], [4, 3, 3, 3, 3, 0, 0, 0]], np.int32)
    nodes = np.zeros((2, 8)); values = np.ones((2, 8)); slopes = np.zeros((2, 8))
    nodes[0, :2] = [1.0, 4.0]
    nodes[1, :4] = [0.5, 1.25, 2.75, 3.5]
    values[0, :4] = a
    values[1, :4] = [b[0], b[1], b[2], b[0]]
    slopes[1, :4] = [d, -d, d / 2, d]
    # the self-test interface carries 4 parameters per call: u1 w.r.t. a0..a3; u2 w.r.t. (b0, b1, b2, d)
    dv, dm = np.zeros((2, 8, 4)), np.zeros((2, 8, 4))
    dv[0, :4, :] = np.eye(4)
    dv[1, 0, 0] = dv[1, 3, 0] = 1.0
    dv[1, 1, 1] = dv[1, 2, 2] = 1.0
    dm[1, :4, 3] = [1.0, -1.0, 0.5, 1.0]

    def u(t):
        val, sens, err = spline_run(cfg, nodes, values, slopes, dv, dm, np.full((2, 1), t))
        assert not err.any()
        return val[:, 0], sens[:, 0, :]

    def rhs(t, z):     # z = x1, x2, then sx[ip][x] for the 9 parameters
        val, sens = u(t)
        x1, x2 = z[0], z[1]
        s = z[2:].reshape(9, 2)
        ds = np.zeros((9, 2))
        du1 = np.zeros(9); du2 = np.zeros(9)
        du1[1:5] = sens[0]
        du2[5:8] = sens[1, :3]
        du2[8] = sens[1, 3]
        ds[:, 0] = -kd * s[:, 0] + du1
        ds[0, 0] -= x1
        ds[:, 1] = -s[:, 1] + du2
        return np.concatenate([[-kd * x1 + val[0], -x2 + val[1]], ds.reshape(-1)])

    z0 = np.zeros(20)
    z0[0] = x10
    # the splines are only C1 (C0 at the extrapolation joints): integrate piecewise between all nodes
    breaks = sorted(set([0.0, 6.0, 1.0, 2.0, 3.0, 4.0, 0.5, 1.25, 2.75, 3.5] + [0.5 + 3.0 * k_ + off for k_ in (-1, 1) for off in (0.0, 0.75, 2.25)]))
    breaks = [t for t in breaks if 0.0 <= t <= 6.0]
    out_t, out_z = [0.0], [z0]
    z = z0
    for t0, t1 in zip(breaks[:-1], breaks[1:]):
        te = [t for t in SPLINE_MODES_TS if t0 < t <= t1]
        sol = solve_ivp(rhs, (t0, t1), z, method="DOP853", rtol=1e-12, atol=1e-14, t_eval=sorted(set(te + [t1])))
        for t, col in zip(sol.t, sol.y.T):
            if any(abs(t - q) < 1e-12 for q in SPLINE_MODES_TS) and t > out_t[-1] + 1e-12:
                out_t.append(t); out_z.append(col)
        z = sol.y[:, -1]
    Z = np.array(out_z)
    assert len(Z) == len(SPLINE_MODES_TS)
    uvals = np.array([u(t)[0][1] for t in SPLINE_MODES_TS])
    return Z[:, :2], Z[:, 2:].reshape(-1, 9, 2), uvals
