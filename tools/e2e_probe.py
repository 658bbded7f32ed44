import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch, amici_b200 as ab
import bench
model = ab.get_model("jakstat_adjoint"); model.set_timepoints(bench.TS)
solver = model.create_solver(); solver.set_relative_tolerance(1e-8); solver.set_absolute_tolerance(1e-12)
solver.set_sensitivity_order(ab.SensitivityOrder.first)
n = 65536
p, k, noise = bench.synthetic_inputs(n, 2, np.array(model.get_free_parameters()), np.array(model.get_fixed_parameters()))
my = 0.5 + noise; sig = np.full_like(my, 0.1)
batch = ab.ExpDataBatch(bench.TS, p, k, my=my, sigmay=sig)
orig = ab._pinned_empty
tal = [0.0]
def timed(shape, dtype):
    t = time.perf_counter(); r = orig(shape, dtype); tal[0] += time.perf_counter() - t; return r
ab._pinned_empty = timed
for chunk in ["8192", "8192", "4096", "16384", "65536"]:
    os.environ["AMICI_B200_CHUNK"] = chunk
    tal[0] = 0.0
    t0 = time.perf_counter(); res = ab.run_simulations(model, solver, batch); t1 = time.perf_counter()
    print(f"chunk {chunk}: total {1e3*(t1-t0):.0f} ms, pinned alloc {1e3*tal[0]:.0f} ms, sum of chunk kernel ms {res.kernel_ms:.0f}", flush=True)
    del res
