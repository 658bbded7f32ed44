#!/usr/bin/env python
"""the reference-vs-reference noise floor regenerated on the ACTUAL bench inputs (first K conditions of
the JAK-STAT batch of bench.py), next to the engine's deviations — for both sensitivity correctors.

Rows: pairs of runs; columns: per field the number of elements that violate |a - b| <= 10 atol + 10 rtol |b| (rtol 1e-8, atol 1e-12)
and the largest relative deviation.  The engine rows come from the host build of the solver source (tests/cpu_emul: same source
as the CUDA kernels, one-lane groups) so that the table can be produced without a GPU; the GPU tests assert the same protocol.

    python tools/noise_floor.py [K] > profiles/r02_noise_floor.md
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from amici_b200 import _capi  # noqa: E402
from oracle.refsim import RefModel  # noqa: E402

RTOL, ATOL = 1e-8, 1e-12
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c = bench.CONFIGS["jakstat"]
ref = RefModel("jakstat_adjoint")
w = c["inputs"](K, c["seed"], ref.p_nominal, ref.k_nominal)
rng = np.random.default_rng(0)
my = rng.uniform(0.1, 1.0, (K, len(c["ts"]), 3))
sig = np.full_like(my, 0.1)
fields = ["status", "x", "sx", "llh", "sllh", "numsteps"]
thr = ref.max_threads()


def run_ref(**kw):
    return ref.run(c["ts"], w["p"], w["k"], my=my, sigmay=sig, fields=fields, num_threads=thr, **kw)


def run_emul(ism):
    lib = C.CDLL(os.path.join(ROOT, "tests", "cpu_emul", "libemul.so"))
    b, out, keep = _capi.make_batch(dict(nx=9, np=17, nk=2, ny=3), c["ts"], w["p"], w["k"], my=my, sigmay=sig)
    o = _capi.default_options(rtol=RTOL, atol=ATOL, sensi=1, ism=ism)
    assert lib.emul_run(b"jakstat_adjoint", C.byref(o), C.byref(b)) == 0
    return out


runs = {"reference KLU, simultaneous": run_ref(linsol=9, ism=1), "reference dense, simultaneous": run_ref(linsol=1, ism=1),
        "reference KLU, staggered": run_ref(linsol=9, ism=2), "reference truth (rtol 1e-12 / atol 1e-16)": run_ref(linsol=9, rtol=1e-12, atol=1e-16, maxsteps=500000),
        "engine, simultaneous": run_emul(1), "engine, staggered": run_emul(2)}
pairs = [("reference dense, simultaneous", "reference KLU, simultaneous"), ("reference KLU, staggered", "reference KLU, simultaneous"),
         ("reference KLU, simultaneous", "reference truth (rtol 1e-12 / atol 1e-16)"), ("reference KLU, staggered", "reference truth (rtol 1e-12 / atol 1e-16)"),
         ("engine, simultaneous", "reference KLU, simultaneous"), ("engine, staggered", "reference KLU, staggered"),
         ("engine, simultaneous", "reference truth (rtol 1e-12 / atol 1e-16)"), ("engine, staggered", "reference truth (rtol 1e-12 / atol 1e-16)")]
print(f"# Round 2 — noise floor on the bench inputs (tools/noise_floor.py, first {K} JAK-STAT conditions of bench.py's batch)\n")
print("Per pair and field: number of elements outside `|a - b| <= 10 atol + 10 rtol |b|` (rtol 1e-8, atol 1e-12) / number of elements, "
      "largest relative deviation.  Status codes are identical in every run; steps per solve: " +
      ", ".join(f"{k_} {v['numsteps'][:, -1].mean():.1f}" for k_, v in runs.items()) + ".\n")
print("| a | b | x | sx | llh | sllh |\n|---|---|---|---|---|---|")
for a, b in pairs:
    cells = []
    for f in ("x", "sx", "llh", "sllh"):
        va, vb = np.asarray(runs[a][f], float), np.asarray(runs[b][f], float).reshape(np.shape(runs[a][f]))
        viol = np.abs(va - vb) > 10 * ATOL + 10 * RTOL * np.abs(vb)
        rel = np.abs(va - vb) / np.maximum(np.abs(vb), 1e-300)
        rel = rel[np.abs(vb) > 1e-10]
        cells.append(f"{int(viol.sum())} / {viol.size} ({rel.max():.1e})")
    print(f"| {a} | {b} | " + " | ".join(cells) + " |")
    assert np.array_equal(runs[a]["status"], runs[b]["status"])
