"""ctypes mirror of the C ABI structs declared in include/amici_b200.h."""
from __future__ import annotations

import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class SolverOptions(C.Structure):
    """``ab_solver_options`` (include/amici_b200.h)."""
    _fields_ = [(n, C.c_double) for n in ("rtol", "atol", "rtol_fsa", "atol_fsa", "t0", "hmax_inv")] + [
        (n, C.c_int) for n in ("sensi", "ism", "maxsteps", "stldet", "max_nonlin_iters", "max_conv_fails",
                               "max_err_test_fails", "reporting")]


# (name, dtype, shape(dims, N, nt, nplist)) for every output buffer of ab_batch
OUT_FIELDS = [
    ("status", "i", lambda d, N, nt, q: (N,)),
    ("llh", "d", lambda d, N, nt, q: (N,)),
    ("chi2", "d", lambda d, N, nt, q: (N,)),
    ("sllh", "d", lambda d, N, nt, q: (N, q)),
    ("x", "d", lambda d, N, nt, q: (N, nt, d["nx"])),
    ("sx", "d", lambda d, N, nt, q: (N, nt, q, d["nx"])),
    ("y", "d", lambda d, N, nt, q: (N, nt, d["ny"])),
    ("sy", "d", lambda d, N, nt, q: (N, nt, q, d["ny"])),
    ("sigmay", "d", lambda d, N, nt, q: (N, nt, d["ny"])),
    ("ssigmay", "d", lambda d, N, nt, q: (N, nt, q, d["ny"])),
    ("res", "d", lambda d, N, nt, q: (N, nt * d["ny"])),
    ("sres", "d", lambda d, N, nt, q: (N, nt * d["ny"], q)),
    ("FIM", "d", lambda d, N, nt, q: (N, q, q)),
    ("x0", "d", lambda d, N, nt, q: (N, d["nx"])),
    ("sx0", "d", lambda d, N, nt, q: (N, q, d["nx"])),
    ("xdot", "d", lambda d, N, nt, q: (N, d["nx"])),
    ("J", "d", lambda d, N, nt, q: (N, d["nx"], d["nx"])),
    ("numsteps", "i", lambda d, N, nt, q: (N, nt)),
    ("numrhsevals", "i", lambda d, N, nt, q: (N, nt)),
    ("numerrtestfails", "i", lambda d, N, nt, q: (N, nt)),
    ("numnonlinsolvconvfails", "i", lambda d, N, nt, q: (N, nt)),
    ("order", "i", lambda d, N, nt, q: (N, nt)),
]


class BatchView(C.Structure):
    """``ab_batch`` (include/amici_b200.h): borrowed host (or device) pointers, row-major."""
    _fields_ = [("N", C.c_int), ("nt", C.c_int), ("nplist", C.c_int),
                ("ts", _dp), ("plist", _ip), ("pscale", _ip), ("p", _dp), ("k", _dp),
                ("my", _dp), ("sigmay_in", _dp)] + [
        (n, _dp if t == "d" else _ip) for n, t, _ in OUT_FIELDS]


def default_options(**kw) -> SolverOptions:
    o = SolverOptions(rtol=1e-8, atol=1e-16, rtol_fsa=float("nan"), atol_fsa=float("nan"), t0=0.0, hmax_inv=0.0,
                      sensi=0, ism=1, maxsteps=10000, stldet=1, max_nonlin_iters=3, max_conv_fails=10,
                      max_err_test_fails=7, reporting=0)
    for k_, v in kw.items():
        setattr(o, k_, v)
    if o.rtol_fsa != o.rtol_fsa:
        o.rtol_fsa = o.rtol
    if o.atol_fsa != o.atol_fsa:
        o.atol_fsa = o.atol
    return o


def make_batch(dims: dict, ts, p, k, plist=None, pscale=None, my=None, sigmay=None, fields=None):
    """Build an ab_batch over freshly allocated numpy outputs. Returns (batch, outputs dict, keepalive)."""
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
    N = p.shape[0]
    k = np.ascontiguousarray(np.atleast_2d(k), dtype=np.float64).reshape(N, dims["nk"])
    plist = np.arange(dims["np"], dtype=np.int32) if plist is None else np.ascontiguousarray(plist, dtype=np.int32)
    pscale = np.zeros(dims["np"], dtype=np.int32) if pscale is None else np.ascontiguousarray(pscale, dtype=np.int32)
    nt, q = len(ts), len(plist)
    b = BatchView(N=N, nt=nt, nplist=q)
    keep = [ts, p, k, plist, pscale]
    b.ts, b.p, b.k = ts.ctypes.data_as(_dp), p.ctypes.data_as(_dp), k.ctypes.data_as(_dp)
    b.plist, b.pscale = plist.ctypes.data_as(_ip), pscale.ctypes.data_as(_ip)
    for name, arr in (("my", my), ("sigmay_in", sigmay)):
        if arr is not None:
            a = np.ascontiguousarray(arr, dtype=np.float64).reshape(N, nt, dims["ny"])
            keep.append(a)
            setattr(b, name, a.ctypes.data_as(_dp))
    out = {}
    for name, t, shp in OUT_FIELDS:
        if fields is not None and name not in fields and name != "status":
            continue
        a = np.zeros(shp(dims, N, nt, q), dtype=np.float64 if t == "d" else np.int32)
        out[name] = a
        setattr(b, name, a.ctypes.data_as(_dp if t == "d" else _ip))
    return b, out, keep
