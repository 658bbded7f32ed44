#!/usr/bin/env python
"""Parity report at BASELINE sizes: the CUDA engine integrates the full synthetic batch of each config; the unmodified
reference (oracle/_ref) integrates the first K conditions of the same batch on the host, with its default linear solver (KLU)
and with the dense one (the reference's own noise floor).  Prints a markdown table of

  max over elements of |a - b| / (10*atol + 10*rtol*|b|)      (1.0 = the tolerance BASELINE.json states)

for engine-vs-KLU and dense-vs-KLU, plus status agreement and step-count statistics.  Run on the GPU box:
  python tools/parity_report.py > gpurun_out/parity_report.md
"""
import sys

import numpy as np

sys.path.insert(0, ".")
import amici_b200 as ab
from oracle.refsim import RefModel

RTOL, ATOL = 1e-8, 1e-12


def scaled(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float).reshape(np.shape(a))
    both = ~(np.isnan(a) | np.isnan(b))
    if not both.any():
        return 0.0
    return float(np.max(np.abs(a[both] - b[both]) / (10 * ATOL + 10 * RTOL * np.abs(b[both]))))


def report(title, model, refname, ts, p, k, K, fields, sensi=True, k_preeq=None, setup=None, ref_kw=None, my=None, batch_kw=None,
           ref_batch_kw=None):
    m = ab.get_model(model)
    m.set_timepoints(ts)
    if setup:
        setup(m)
    s = m.create_solver()
    s.set_relative_tolerance(RTOL)
    s.set_absolute_tolerance(ATOL)
    if sensi:
        s.set_sensitivity_order(ab.SensitivityOrder.first)
    if k_preeq is not None:
        s.set_newton_max_steps(50)
    N = len(p)
    sig = None if my is None else np.full_like(my, 0.1)
    want = ["status", "numsteps"] + fields
    res = ab.run_batch(m, s, ab.ExpDataBatch(ts, p, k, my=my, sigmay=sig, k_preeq=k_preeq, **(batch_kw or {})), fields=want)
    ref = RefModel(refname)
    kw = dict(sensi_order=1 if sensi else 0, rtol=RTOL, atol=ATOL, num_threads=ref.max_threads(), fields=want, **(ref_kw or {}))
    sl = slice(0, K)
    common = dict(my=None if my is None else my[sl], sigmay=None if sig is None else sig[sl], k_preeq=None if k_preeq is None else k_preeq[sl])
    common.update({k_: (v[sl] if isinstance(v, np.ndarray) else v) for k_, v in (ref_batch_kw or {}).items()})
    r_klu = ref.run(ts, p[sl], k[sl], linsol=9, **common, **kw)
    r_dense = ref.run(ts, p[sl], k[sl], linsol=1, **common, **kw)
    st_g, st_r = res["status"][sl], r_klu["status"]
    print(f"\n### {title}\n")
    print(f"{N} conditions on the GPU ({int((res['status'] != 0).sum())} with non-zero status), first {K} also by the reference; "
          f"status identical on {int((st_g == st_r).sum())} / {K} (reference non-zero: {int((st_r != 0).sum())}); "
          f"steps per solve: engine {res['numsteps'][sl, -1].mean():.1f} +- {res['numsteps'][sl, -1].std():.1f}, "
          f"reference {r_klu['numsteps'][:, -1].mean():.1f} +- {r_klu['numsteps'][:, -1].std():.1f}.\n")
    print("| field | engine vs reference (KLU) | reference dense vs KLU |\n|---|---|---|")
    ok = (st_g == 0) & (st_r == 0) & (r_dense["status"] == 0)
    for f in fields:
        a, b, c = np.asarray(res[f])[sl][ok], np.asarray(r_klu[f]).reshape(np.asarray(res[f])[sl].shape)[ok], np.asarray(r_dense[f]).reshape(np.asarray(res[f])[sl].shape)[ok]
        print(f"| {f} | {scaled(a, b):.3g} | {scaled(c, b):.3g} |")


def nominal(model):
    m = ab.get_model(model)
    return np.array(m.get_free_parameters()), np.array(m.get_fixed_parameters())


def noisy_measurements(refname, ts, p, k, rng, **kw):
    """measurements = noise-free observables of the reference-free engine run + N(0, 0.1^2) (bench.py does the same)."""
    return None


def presim_report(N=65536, K=2048):
    """configs[2]-sized steadystate batch with the full condition context: pre-equilibration, presimulation with its own fixed
    parameters (a quarter of the conditions have none), re-initialisation of states 0 and 2 for the simulation."""
    p0, k0 = nominal("steadystate")
    rng = np.random.default_rng(5)
    p = np.tile(p0, (N, 1)) * 10 ** rng.normal(0, 0.1, (N, 5))
    k, kp, kps = (np.tile(k0, (N, 1)) * rng.uniform(0.5, 2, (N, 4)) for _ in range(3))
    tps = rng.uniform(0.1, 5.0, N)
    tps[::4] = 0.0
    report("steadystate: pre-equilibration + presimulation + state re-initialisation, forward sensitivities", "steadystate", "steadystate",
           np.linspace(0, 10, 11), p, k, K, ["x0", "sx0", "x", "sx", "llh", "sllh", "x_ss", "sx_ss"], k_preeq=kp,
           my=rng.uniform(0.1, 1.0, (N, 11, 3)),
           setup=lambda m: (m.set_steady_state_computation_mode(2), m.set_steady_state_sensitivity_mode(0)),
           ref_kw=dict(newton_maxsteps=50, ss_mode=2, ss_sensi_mode=0, reinit_states_sim=0b101),
           batch_kw=dict(k_presim=kps, t_presim=tps, reinit_states_sim=[0, 2]), ref_batch_kw=dict(k_presim=kps, t_presim=tps))


if __name__ == "__main__":
    if sys.argv[1:] == ["presim"]:
        presim_report()
        sys.exit(0)
    print("# Round 1 — parity at BASELINE sizes (tools/parity_report.py, run on the B200 box)\n")
    print("Unit: multiples of the tolerance BASELINE.json states (10*rtol relative / 10*atol absolute at rtol 1e-8, atol 1e-12); the second "
          "column is the reference's own spread between its two linear solvers on the same inputs.")
    N, K = 65536, 2048
    # configs[1]
    p0, k0 = nominal("jakstat_adjoint")
    rng = np.random.default_rng(2)
    p = np.tile(p0, (N, 1)) * 10 ** rng.normal(0, 0.1, (N, 17))
    k = np.tile(k0, (N, 1)) * rng.uniform(0.8, 1.25, (N, 2))
    ts = np.arange(0, 61, 5.0)
    my = rng.uniform(0.0, 1.0, (N, len(ts), 3))
    report("configs[1] JAK-STAT, 17 forward sensitivities", "jakstat_adjoint", "jakstat_adjoint", ts, p, k, K,
           ["x", "y", "llh", "sllh", "sx"], my=my)
    # configs[0]
    p0, k0 = nominal("robertson_ode")
    rng = np.random.default_rng(1)
    n0 = 4096
    tsr = np.concatenate([[0.0], 4 * np.logspace(-6, 6, 50)])
    report("configs[0] Robertson (ODE form on the GPU, the shipped DAE model through IDAS as reference), 4096 log-uniform draws",
           "robertson_ode", "robertson", tsr, np.tile(p0, (n0, 1)) * 10 ** rng.uniform(-1, 1, (n0, 3)), np.full((n0, 1), 0.9), 1024,
           ["x", "y"], sensi=False)
    # configs[2]
    p0, k0 = nominal("steadystate")
    rng = np.random.default_rng(3)
    p = np.tile(p0, (N, 1)) * 10 ** rng.normal(0, 0.1, (N, 5))
    k = np.tile(k0, (N, 1)) * rng.uniform(0.5, 2, (N, 4))
    kp = np.tile(k0, (N, 1)) * rng.uniform(0.5, 2, (N, 4))
    tss = np.linspace(0, 10, 11)
    report("configs[2] steadystate, pre-equilibration (Newton, simulation if it fails) + forward sensitivities", "steadystate", "steadystate",
           tss, p, k, K, ["x", "sx", "llh", "sllh", "x_ss", "sx_ss"], k_preeq=kp, my=rng.uniform(0.1, 1.0, (N, 11, 3)),
           setup=lambda m: (m.set_steady_state_computation_mode(2), m.set_steady_state_sensitivity_mode(0)),
           ref_kw=dict(newton_maxsteps=50, ss_mode=2, ss_sensi_mode=0))
    # configs[3]
    for model, tsm in (("events", np.linspace(0, 10, 21)), ("neuron", np.linspace(0, 100, 21))):
        p0, k0 = nominal(model)
        rng = np.random.default_rng(4)
        report(f"configs[3] {model}: events, root finding, forward sensitivities", model, model, tsm,
               np.tile(p0, (N, 1)) * 10 ** rng.normal(0, 0.1, (N, len(p0))), np.tile(k0, (N, 1)), 1024 if model == "neuron" else K,
               ["x", "y", "sx", "z", "sz"])
