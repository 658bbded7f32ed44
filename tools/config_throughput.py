#!/usr/bin/env python
"""Device-resident throughput of the BASELINE.json configs other than the headline one (bench.py = configs[1]):
configs[0] Robertson (4096 log-uniform draws, no sensitivities), configs[2] steadystate with pre-equilibration,
configs[3] events / neuron, all with the synthetic inputs of SURVEY.md §8d.  One JSON line per config."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import amici_b200 as ab
from amici_b200.device import DeviceBatch


def run(name, model, ts, p, k, sensi=True, k_preeq=None, setup=None, reps=3):
    m = ab.get_model(model)
    m.set_timepoints(ts)
    if setup:
        setup(m)
    s = m.create_solver()
    s.set_relative_tolerance(1e-8)
    s.set_absolute_tolerance(1e-12)
    if sensi:
        s.set_sensitivity_order(ab.SensitivityOrder.first)
    if k_preeq is not None:
        s.set_newton_max_steps(50)
    s.set_return_data_reporting_mode(ab.RDataReporting.likelihood)
    N = len(p)
    rng = np.random.default_rng(0)
    my = rng.uniform(0.1, 1.0, (N, len(ts), m.ny))
    batch = ab.ExpDataBatch(ts, p, k, my=my, sigmay=np.full_like(my, 0.1), k_preeq=k_preeq)
    db = DeviceBatch(m, s, batch, torch.device("cuda", 0))
    db.solve()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        db.solve()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    st = db.out["status"].cpu().numpy()
    print(json.dumps({"config": name, "model": model, "n_conditions": N, "sensitivities": bool(sensi), "ms_per_pass": ms,
                      "solves_per_s": N / (ms * 1e-3), "n_failed": int((st != 0).sum()), "mean_steps": float(db.out["numsteps"][:, -1].double().mean())}))


def nominal(model):
    m = ab.get_model(model)
    return np.array(m.get_free_parameters()), np.array(m.get_fixed_parameters())


if __name__ == "__main__":
    N = 65536
    p0, k0 = nominal("robertson_ode")
    rng = np.random.default_rng(1)
    ts = np.concatenate([[0.0], 4 * np.logspace(-6, 6, 50)])
    run("configs[0] Robertson, 4096 log-uniform draws, forward simulation", "robertson_ode", ts,
        np.tile(p0, (4096, 1)) * 10 ** rng.uniform(-1, 1, (4096, 3)), np.full((4096, 1), 0.9), sensi=False)
    nbig = 262144   # the 4096-draw batch of configs[0] occupies 128 warps of the chip's ~9,500: same workload, 64x the draws
    run("configs[0] workload at 262,144 draws (one thread per condition fills the chip)", "robertson_ode", ts,
        np.tile(p0, (nbig, 1)) * 10 ** rng.uniform(-1, 1, (nbig, 3)), np.full((nbig, 1), 0.9), sensi=False)
    p0, k0 = nominal("steadystate")
    rng = np.random.default_rng(3)
    p = np.tile(p0, (N, 1)) * 10 ** rng.normal(0, 0.1, (N, 5))
    run("configs[2] steadystate, pre-equilibration (Newton, then simulation) + FSA", "steadystate", np.linspace(0, 10, 11), p,
        np.tile(k0, (N, 1)) * rng.uniform(0.5, 2, (N, 4)), k_preeq=np.tile(k0, (N, 1)) * rng.uniform(0.5, 2, (N, 4)),
        setup=lambda m: (m.set_steady_state_computation_mode(2), m.set_steady_state_sensitivity_mode(0)))
    for model, tsm in (("events", np.linspace(0, 10, 21)), ("neuron", np.linspace(0, 100, 21))):
        p0, k0 = nominal(model)
        rng = np.random.default_rng(4)
        run(f"configs[3] {model}, events + root finding + FSA", model, tsm, np.tile(p0, (N, 1)) * 10 ** rng.normal(0, 0.1, (N, len(p0))),
            np.tile(k0, (N, 1)))
