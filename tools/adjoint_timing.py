"""Adjoint vs forward sensitivities on the bench batch (JAK-STAT, 65,536 conditions): kernel times through run_simulations."""
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amici_b200 as ab

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = ab.get_model("jakstat_adjoint")
ts = np.arange(0, 61, 5.0)
m.set_timepoints(ts)
rng = np.random.default_rng(0)
p = np.tile(m.get_free_parameters(), (N, 1)) * 10 ** rng.normal(0, 0.1, (N, m.np()))
k = np.tile(m.get_fixed_parameters(), (N, 1))
s0 = m.create_solver()
s0.set_relative_tolerance(1e-8)
s0.set_absolute_tolerance(1e-12)
y0 = ab.run_simulations(m, s0, ab.ExpDataBatch(ts, p, k))["y"]
my = y0 + 0.1 * rng.standard_normal(y0.shape)
batch = ab.ExpDataBatch(ts, p, k, my=my, sigmay=np.full_like(my, 0.1))
out = {}
for name, meth in (("forward", ab.SensitivityMethod.forward), ("adjoint", ab.SensitivityMethod.adjoint)):
    s = m.create_solver()
    s.set_relative_tolerance(1e-8)
    s.set_absolute_tolerance(1e-12)
    s.set_sensitivity_order(ab.SensitivityOrder.first)
    s.set_sensitivity_method(meth)
    s.set_return_data_reporting_mode(ab.RDataReporting.likelihood)
    best = None
    for rep in range(3):
        r = ab.run_simulations(m, s, batch)
        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
    out[name] = dict(kernel_ms=best, solves_per_s=N / best * 1e3, n_failed=int((r["status"] != 0).sum()), sum_llh=float(r["llh"].sum()),
                     sum_sllh0=float(r["sllh"][:, 0].sum()))
print(json.dumps(out))
