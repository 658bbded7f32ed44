#!/usr/bin/env python
"""Do two library variants take the same steps?  Runs the JAK-STAT bench inputs (first N conditions) through each and compares
status / step counters / llh condition by condition:   python tools/ab_diff.py 4096 _base ""  """
import os
import subprocess
import sys

import numpy as np

if sys.argv[1] != "--one":
    n = sys.argv[1]
    for v in sys.argv[2:]:
        env = dict(os.environ, AMICI_B200_LIB=os.path.abspath(f"amici_b200/lib/libamici_b200{v}.so"))
        subprocess.run([sys.executable, __file__, "--one", n, f"/tmp/ab_diff{v or '_default'}.npz"], env=env, check=True)
    a = np.load(f"/tmp/ab_diff{sys.argv[2] or '_default'}.npz")
    for v in sys.argv[3:]:
        b = np.load(f"/tmp/ab_diff{v or '_default'}.npz")
        for f in a.files:
            same = np.array_equal(a[f], b[f], equal_nan=True)
            msg = f"{sys.argv[2] or 'default'} vs {v or 'default'}: {f:24s} identical={same}"
            if not same:
                diff = (a[f] != b[f]).reshape(len(a[f]), -1).any(1)
                msg += f"  conditions differing: {int(diff.sum())} / {len(diff)}  first: {np.flatnonzero(diff)[:5]}"
            print(msg)
    sys.exit(0)
sys.path.insert(0, ".")
import amici_b200 as ab
import bench
c = bench.CONFIGS["jakstat"]
n = int(sys.argv[2])
model, make_solver, w = bench.build_engine_objects(ab, c, n, 0)
rng = np.random.default_rng(0)
my = rng.uniform(0.1, 1.0, (n, len(c["ts"]), 3))
res = ab.run_batch(model, make_solver(True), ab.ExpDataBatch(c["ts"], w["p"], w["k"], my=my, sigmay=np.full_like(my, 0.1)),
                   fields=["status", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails", "order", "llh", "sllh", "x"])
np.savez(sys.argv[3], **{f: np.asarray(res[f]) for f in ("status", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails", "order", "llh", "x")})
