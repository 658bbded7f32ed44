import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from tests.test_gpu import _run, GOLD
for fx in ("events_presim", "events_preeq", "events_preeq_presim", "neuron_presim"):
    g = np.load(os.path.join(GOLD, fx + ".npz"))
    res = _run(fx.split("_")[0], g)
    np.savez("gpurun_out/dump_%s.npz" % fx, **{f: np.asarray(res[f]) for f in ("status", "x", "sx", "y", "llh", "chi2", "sllh", "x0", "sx0", "z", "sz", "numsteps", "res")})
