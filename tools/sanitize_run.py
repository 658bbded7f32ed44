"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import amici_b200 as ab

def run(model, N, ts, sensi=True, **kw):
    m = ab.get_model(model)
    m.set_timepoints(ts)
    s = m.create_solver()
    s.set_relative_tolerance(1e-6)
    s.set_absolute_tolerance(1e-10)
    if sensi:
        s.set_sensitivity_order(ab.SensitivityOrder.first)
    rng = np.random.default_rng(0)
    p = np.tile(m.get_free_parameters(), (N, 1)) * 10 ** rng.normal(0, 0.05, (N, m.np()))
    k = np.tile(m.get_fixed_parameters(), (N, 1))
    my = rng.uniform(0.2, 1.0, (N, len(ts), m.ny))
    res = ab.run_simulations(m, s, ab.ExpDataBatch(ts, p, k, my=my, sigmay=np.full_like(my, 0.1), **kw))
    print(model, N, "status ok:", int((res["status"] == 0).sum()), "of", N)

which = sys.argv[1:] or ["jakstat", "steadystate", "presim", "events", "massaction_small"]
if "jakstat" in which:
    run("jakstat_adjoint", 20, np.arange(0, 61, 10.0))
if "steadystate" in which:
    run("steadystate", 40, np.linspace(0, 10, 6), k_preeq=np.tile([0.1, 0.4, 0.7, 1.0], (40, 1)) * 1.3)
if "presim" in which:   # presimulation + state re-initialisation + post-equilibration, after a pre-equilibration and alone
    kp = np.tile([0.1, 0.4, 0.7, 1.0], (40, 1))
    run("steadystate", 40, np.concatenate([np.linspace(0, 10, 6), [np.inf]]), k_preeq=kp * 1.3, k_presim=kp * 0.8,
        t_presim=np.linspace(0, 4, 40), reinit_states_sim=[0, 2], reinit_states_presim=[1])
    run("jakstat_adjoint", 12, np.arange(0, 61, 10.0), t_presim=np.linspace(0, 6, 12))
if "events" in which:
    run("events", 40, np.linspace(0, 10, 11))
if "massaction_small" in which:
    run("massaction_small", 4, np.linspace(0, 10, 6))
