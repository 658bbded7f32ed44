"""Adjoint vs forward-sensitivity gradient of the engine on a seeded JAK-STAT batch; dumps the worst conditions (diagnostic)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amici_b200 as ab

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = ab.get_model("jakstat_adjoint")
ts = np.arange(0, 61, 5.0)
m.set_timepoints(ts)
rng = np.random.default_rng(51)
p = np.tile(m.get_free_parameters(), (N, 1)) * 10 ** rng.normal(0, 0.1, (N, m.np()))
k = np.tile(m.get_fixed_parameters(), (N, 1))
s0 = m.create_solver()
s0.set_relative_tolerance(1e-8)
s0.set_absolute_tolerance(1e-12)
y0 = ab.run_simulations(m, s0, ab.ExpDataBatch(ts, p, k))["y"]
my = y0 + 0.1 * rng.standard_normal(y0.shape)
batch = ab.ExpDataBatch(ts, p, k, my=my, sigmay=np.full_like(my, 0.1))
res = {}
for name, meth in (("fsa", ab.SensitivityMethod.forward), ("asa", ab.SensitivityMethod.adjoint)):
    s = m.create_solver()
    s.set_relative_tolerance(1e-8)
    s.set_absolute_tolerance(1e-12)
    s.set_sensitivity_order(ab.SensitivityOrder.first)
    s.set_sensitivity_method(meth)
    res[name] = ab.run_simulations(m, s, batch)
a, f = res["asa"]["sllh"], res["fsa"]["sllh"]
scale = np.abs(f).max(axis=1, keepdims=True)
d = (np.abs(a - f) / (1e-3 * np.abs(f) + 1e-5 * scale)).max(axis=1)
order = np.argsort(-d)[:16]
print("worst", d[order][:8], order[:8], "n over", int((d > 1).sum()))
np.savez("gpurun_out/adjoint_cross.npz", idx=order, p=p[order], k=k[order], my=my[order], asa=a[order], fsa=f[order],
         nst_asa=res["asa"]["numsteps"][order], nst_fsa=res["fsa"]["numsteps"][order], d=d[order], llh_asa=res["asa"]["llh"][order], llh_fsa=res["fsa"]["llh"][order])
