"""Throughput of the block solver on the mass-action network: python tools/ma_throughput.py <model> <N> [nplist]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import amici_b200 as ab
from amici_b200.codegen.massaction import network
name, N = sys.argv[1], int(sys.argv[2])
net = network(name)
npl = int(sys.argv[3]) if len(sys.argv) > 3 else net["nr"]
model = ab.get_model(name)
ts = np.linspace(0, 10, 11)
model.set_timepoints(ts)
model.set_parameter_list(list(range(npl)))
solver = model.create_solver()
solver.set_relative_tolerance(1e-8); solver.set_absolute_tolerance(1e-12)
solver.set_sensitivity_order(ab.SensitivityOrder.first)
rng = np.random.default_rng(5)
p = np.tile(net["p_nominal"], (N, 1)) * 10 ** rng.uniform(-1, 1, (N, net["nr"]))
k = np.ones((N, 1))
my = rng.uniform(0.5, 1.5, (N, 11, net["ny"])); sig = np.full_like(my, 0.1)
batch = ab.ExpDataBatch(ts, p, k, my=my, sigmay=sig)
for rep in range(2):
    t0 = time.time()
    res = ab.run_batch(model, solver, batch, fields=["status", "llh", "sllh", "numsteps"])
    t1 = time.time()
    print(f"{name} N={N} nplist={npl}: wall {t1-t0:.3f}s kernel {res.kernel_ms:.1f} ms -> {N/(res.kernel_ms*1e-3):.1f} solves/s; failed {(res['status']!=0).sum()}, mean steps {res['numsteps'][:,-1].mean():.0f}")
