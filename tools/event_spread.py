#!/usr/bin/env python
"""Where do the event models' sensitivities sit relative to the reference's own spread?  (VERDICT r01 "weak" 2)

For the first K draws of BASELINE config 4 (events / neuron, p x 10^N(0, 0.1^2), seed 4) this prints per-condition error
statistics, in units of 10*atol + 10*rtol*|truth|, of
   engine vs truth | reference KLU vs truth | reference dense vs truth      and      engine vs KLU | dense vs KLU
as quantiles over conditions (max over the elements of a condition), so that one outlier condition does not decide the
comparison.  AMICI_B200_LIB selects an engine build (A/B: exact divisions / exact pow in the step control).
  python tools/event_spread.py neuron 1024
"""
import sys

import numpy as np

sys.path.insert(0, ".")
import amici_b200 as ab
from oracle.refsim import RefModel

RTOL, ATOL = 1e-8, 1e-12
model, K = sys.argv[1], int(sys.argv[2])
ts = np.linspace(0, 10, 21) if model == "events" else np.linspace(0, 100, 21)
nme = 2 if model == "events" else 10
m = ab.get_model(model)
m.set_timepoints(ts)
m.set_n_max_event(nme)
s = m.create_solver()
s.set_relative_tolerance(RTOL)
s.set_absolute_tolerance(ATOL)
s.set_sensitivity_order(ab.SensitivityOrder.first)
rng = np.random.default_rng(4)
p = np.tile(m.get_free_parameters(), (65536, 1)) * 10 ** rng.normal(0, 0.1, (65536, m.np()))
p, k = p[:K], np.tile(m.get_fixed_parameters(), (K, 1))
fields = ["x", "sx", "z", "sz"]
res = ab.run_batch(m, s, ab.ExpDataBatch(ts, p, k), fields=["status", "numsteps"] + fields)
ref = RefModel(model)
kw = dict(nmaxevent=nme, num_threads=ref.max_threads(), fields=["status", "numsteps"] + fields)
klu = ref.run(ts, p, k, linsol=9, **kw)
dense = ref.run(ts, p, k, linsol=1, **kw)
truth = ref.run(ts, p, k, linsol=9, rtol=1e-12, atol=1e-16, maxsteps=2000000, **kw)
ok = (res["status"] == 0) & (klu["status"] == 0) & (dense["status"] == 0) & (truth["status"] == 0)
print(f"## {model}: {K} draws, {int(ok.sum())} succeed everywhere; status identical to KLU on {int((res['status'] == klu['status']).sum())}; "
      f"steps engine {res['numsteps'][ok, -1].mean():.1f} klu {klu['numsteps'][ok, -1].mean():.1f} dense {dense['numsteps'][ok, -1].mean():.1f}")


def per_cond(a, b, t):
    a, b, t = (np.asarray(v, float).reshape(K, -1)[ok] for v in (a, b, t))
    e = np.abs(a - b) / (10 * ATOL + 10 * RTOL * np.abs(t))
    e[~np.isfinite(e)] = 0.0
    return e.max(axis=1)


print("| field | statistic | engine-truth | klu-truth | dense-truth | engine-klu | dense-klu |\n|---|---|---|---|---|---|---|")
for f in fields:
    cols = [per_cond(res[f], truth[f], truth[f]), per_cond(klu[f], truth[f], truth[f]), per_cond(dense[f], truth[f], truth[f]),
            per_cond(res[f], klu[f], truth[f]), per_cond(dense[f], klu[f], truth[f])]
    for name, fn in (("median", np.median), ("p90", lambda v: np.quantile(v, 0.9)), ("p99", lambda v: np.quantile(v, 0.99)), ("max", np.max)):
        print(f"| {f} | {name} | " + " | ".join(f"{fn(c):.3g}" for c in cols) + " |")
