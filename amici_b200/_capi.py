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
                               "max_err_test_fails", "reporting")] + [
        (n, C.c_double) for n in ("rtol_ss", "atol_ss", "rtol_ss_sensi", "atol_ss_sensi", "newton_damping_lower_bound")] + [
        (n, C.c_int) for n in ("newton_maxsteps", "newton_damping_mode", "sensi_ss_check", "ss_mode", "ss_sensi_mode",
                               "nmaxevent", "posteq")] + [
        (n, C.c_uint64) for n in ("reinit_states_sim", "reinit_states_presim", "x_nonneg")] + [("maxtime", C.c_double),
                                                                                               ("sensi_meth", C.c_int), ("reserved_", C.c_int)]


# (name, dtype, shape(dims, N, nt, nplist)) for every output buffer of ab_batch; dims["ne_"] = nmaxevent of the call
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
    ("x_ss", "d", lambda d, N, nt, q: (N, d["nx"])),
    ("sx_ss", "d", lambda d, N, nt, q: (N, q, d["nx"])),
    ("preeq_status", "i", lambda d, N, nt, q: (N, 3)),
    ("preeq_numsteps", "i", lambda d, N, nt, q: (N, 3)),
    ("preeq_wrms", "d", lambda d, N, nt, q: (N,)),
    ("preeq_t", "d", lambda d, N, nt, q: (N,)),
    ("posteq_status", "i", lambda d, N, nt, q: (N, 3)),
    ("posteq_numsteps", "i", lambda d, N, nt, q: (N, 3)),
    ("posteq_wrms", "d", lambda d, N, nt, q: (N,)),
    ("posteq_t", "d", lambda d, N, nt, q: (N,)),
    ("z", "d", lambda d, N, nt, q: (N, d.get("nmaxevent", 10), d.get("nz", 0))),
    ("sz", "d", lambda d, N, nt, q: (N, d.get("nmaxevent", 10), q, d.get("nz", 0))),
    ("rz", "d", lambda d, N, nt, q: (N, d.get("nmaxevent", 10), d.get("nz", 0))),
    ("srz", "d", lambda d, N, nt, q: (N, d.get("nmaxevent", 10), q, d.get("nz", 0))),
    ("sigmaz", "d", lambda d, N, nt, q: (N, d.get("nmaxevent", 10), d.get("nz", 0))),
]
PREEQ_FIELDS = ("x_ss", "sx_ss", "preeq_status", "preeq_numsteps", "preeq_wrms", "preeq_t")
POSTEQ_FIELDS = ("posteq_status", "posteq_numsteps", "posteq_wrms", "posteq_t")
EVENT_FIELDS = ("z", "sz", "rz", "srz", "sigmaz")


class BatchView(C.Structure):
    """``ab_batch`` (include/amici_b200.h): borrowed host (or device) pointers, row-major."""
    _fields_ = [("N", C.c_int), ("nt", C.c_int), ("nplist", C.c_int),
                ("ts", _dp), ("plist", _ip), ("pscale", _ip), ("p", _dp), ("k", _dp),
                ("my", _dp), ("sigmay_in", _dp), ("k_preeq", _dp), ("mz", _dp), ("sigmaz_in", _dp),
                ("x0_in", _dp), ("sx0_in", _dp), ("k_presim", _dp), ("t_presim", _dp)] + [
        (n, _dp if t == "d" else _ip) for n, t, _ in OUT_FIELDS]


def default_options(**kw) -> SolverOptions:
    o = SolverOptions(rtol=1e-8, atol=1e-16, rtol_fsa=float("nan"), atol_fsa=float("nan"), t0=0.0, hmax_inv=0.0,
                      sensi=0, ism=1, maxsteps=10000, stldet=1, max_nonlin_iters=3, max_conv_fails=10,
                      max_err_test_fails=7, reporting=0, rtol_ss=float("nan"), atol_ss=float("nan"),
                      rtol_ss_sensi=float("nan"), atol_ss_sensi=float("nan"), newton_damping_lower_bound=1e-8,
                      newton_maxsteps=0, newton_damping_mode=1, sensi_ss_check=1, ss_mode=2, ss_sensi_mode=2, nmaxevent=10,
                      posteq=0, reinit_states_sim=0, reinit_states_presim=0, x_nonneg=0, maxtime=0.0)
    for k_, v in kw.items():
        setattr(o, k_, v)
    if o.rtol_fsa != o.rtol_fsa:
        o.rtol_fsa = o.rtol
    if o.atol_fsa != o.atol_fsa:
        o.atol_fsa = o.atol
    # steady-state tolerances default to 100x the integration tolerances (reference src/solver.cpp:981-1037)
    for name, base in (("rtol_ss", o.rtol), ("atol_ss", o.atol), ("rtol_ss_sensi", o.rtol), ("atol_ss_sensi", o.atol)):
        if getattr(o, name) != getattr(o, name):
            setattr(o, name, 100.0 * base)
    return o


def make_batch(dims: dict, ts, p, k, plist=None, pscale=None, my=None, sigmay=None, fields=None, k_preeq=None,
               mz=None, sigmaz=None, nmaxevent=10, x0=None, sx0=None, k_presim=None, t_presim=None):
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
    dims = dict(dims, nmaxevent=nmaxevent)
    if k_preeq is not None:
        a = np.ascontiguousarray(k_preeq, dtype=np.float64).reshape(N, dims["nk"])
        keep.append(a)
        b.k_preeq = a.ctypes.data_as(_dp)
    for name, arr in (("mz", mz), ("sigmaz_in", sigmaz)):
        if arr is not None:
            a = np.ascontiguousarray(arr, dtype=np.float64).reshape(N, nmaxevent, dims.get("nz", 0))
            keep.append(a)
            setattr(b, name, a.ctypes.data_as(_dp))
    for name, arr, shape in (("x0_in", x0, (N, dims["nx"])), ("sx0_in", sx0, (N, q, dims["nx"])),
                             ("k_presim", k_presim, (N, dims["nk"])), ("t_presim", t_presim, (N,))):
        if arr is not None:
            a = np.ascontiguousarray(arr, dtype=np.float64).reshape(shape)
            keep.append(a)
            setattr(b, name, a.ctypes.data_as(_dp))
    out = {}
    for name, t, shp in OUT_FIELDS:
        if fields is not None and name not in fields and name != "status":
            continue
        if name in PREEQ_FIELDS and k_preeq is None and (fields is None or name not in fields):
            continue
        if name in EVENT_FIELDS and dims.get("nz", 0) == 0:
            continue
        if name in POSTEQ_FIELDS and np.all(np.isfinite(ts)) and (fields is None or name not in fields):
            continue
        a = np.zeros(shp(dims, N, nt, q), dtype=np.float64 if t == "d" else np.int32)
        out[name] = a
        setattr(b, name, a.ctypes.data_as(_dp if t == "d" else _ip))
    return b, out, keep
