"""Harvest real inputs of CVODES' stability-limit detection (cvSLdet) as test vectors: the host build of the solver source
(tests/cpu_emul, compiled with -DAB_SLDET_RECORD) integrates the golden batches of the stiff models and records the ring
buffer, its head, the order and the result of every cvSLdet call.  A subset that covers every return code that occurs is
written to tests/golden/sldet_cases.npz; the GPU test feeds it to the engine's device forms of the routine.

  python tests/golden/make_sldet_cases.py
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from amici_b200 import _capi  # noqa: E402
from tests.test_cpu import MODEL_DIMS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = "/tmp/libemul_rec.so"
subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DAB_SLDET_RECORD", "-o", lib_path,
                os.path.join(ROOT, "tests", "cpu_emul", "emul.cpp")], check=True)
lib = C.CDLL(lib_path)
lib.emul_sldet_record.argtypes = [C.c_void_p, C.c_long]
lib.emul_sldet_recorded.restype = C.c_long
cap = 4_000_000
buf = np.zeros((cap, 43))
lib.emul_sldet_record(buf.ctypes.data, cap)
rng = np.random.default_rng(0)
for fixture, model, sensi in (("jakstat_adjoint", "jakstat_adjoint", 1), ("robertson_ode", "robertson_ode", 1), ("robertson_ode", "robertson_ode", 0),
                              ("steadystate", "steadystate", 1), ("neuron", "neuron", 1), ("events", "events", 1)):
    g = np.load(os.path.join(HERE, fixture + ".npz"))
    kw = dict(my=g["my"], sigmay=g["sigmay_in"]) if "my" in g else {}
    b, out, keep = _capi.make_batch(MODEL_DIMS[model], g["ts"], g["p"], g["k"], **kw)
    o = _capi.default_options(rtol=1e-8, atol=1e-12, sensi=sensi)
    assert lib.emul_run(model.encode(), C.byref(o), C.byref(b)) == 0
    print(fixture, sensi, "calls so far", lib.emul_sldet_recorded())
n = lib.emul_sldet_recorded()
rec = buf[:n]
codes = rec[:, 42].astype(int)
print({c: int((codes == c).sum()) for c in np.unique(codes)})
pick = []
for c in np.unique(codes):
    idx = np.flatnonzero(codes == c)
    pick.append(rng.choice(idx, min(len(idx), 400), replace=False))
pick = np.sort(np.concatenate(pick))
np.savez_compressed(os.path.join(HERE, "sldet_cases.npz"), ssdat=rec[pick, :40].reshape(-1, 5, 8), head=rec[pick, 40].astype(np.int32),
                    q=rec[pick, 41].astype(np.int32), result=codes[pick].astype(np.int32))
print("wrote", len(pick), "cases", os.path.getsize(os.path.join(HERE, "sldet_cases.npz")), "bytes")
