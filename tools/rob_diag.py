#!/usr/bin/env python
"""Step-count statistics of the Robertson config (4096 log-uniform draws, no sensitivities) for library variants:
   python tools/rob_diag.py "" _rob_ed _rob_er ...   (AMICI_B200_LIB is read by amici_b200.load_library)"""
import os
import subprocess
import sys

import numpy as np

if len(sys.argv) > 1 and sys.argv[1] != "--one":
    for v in sys.argv[1:]:
        env = dict(os.environ, AMICI_B200_LIB=os.path.abspath(f"amici_b200/lib/libamici_b200{v}.so"))
        print("variant=%-10s" % (v or "default"), subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True).stdout.strip())
    sys.exit(0)
sys.path.insert(0, ".")
import amici_b200 as ab
import bench
c = bench.CONFIGS["robertson"]
model, make_solver, w = bench.build_engine_objects(ab, c, 4096, 0)
res = ab.run_batch(model, make_solver(False), ab.ExpDataBatch(c["ts"], w["p"], w["k"]),
                   fields=["status", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails", "x"])
print("nst %.2f nfe %.2f netf %.3f ncfn %.4f failed %d" % tuple([res[f][:, -1].mean() for f in ("numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails")] + [(res["status"] != 0).sum()]),
      "per-tp", np.round(res["numsteps"].mean(0)[::5]))
