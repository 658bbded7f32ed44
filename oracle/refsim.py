"""oracle/refsim.py — TEST INFRASTRUCTURE (checker), not product code.

ctypes binding of oracle/_ref/libref_<model>.so, i.e. of the UNMODIFIED reference
`amici::run_simulations` (reference src/amici.cpp:220-264) compiled by
oracle/build_ref.sh.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class RefDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "nx", "nx_solver", "np", "nk", "ny", "nytrue", "nz", "nztrue", "ne", "ne_solver",
        "nw", "nspl", "is_ode")]


class RefOptions(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "rtol", "atol", "rtol_fsa", "atol_fsa", "ss_tol_factor", "rtol_ss", "atol_ss", "t0")] + [
        (n, C.c_int) for n in (
            "sensi_order", "sensi_meth", "ism", "linsol", "maxsteps", "newton_maxsteps",
            "ss_mode", "ss_sensi_mode", "nmaxevent", "stldet", "reporting", "num_threads",
            "failfast", "reinit_all", "reinit_states_sim", "reinit_states_presim", "x_nonneg")] + [("maxtime", C.c_double)]


_IN_D = ("ts",)
_OUT_FIELDS = [  # name, ctype, shape-lambda(d, N, nt, npl, nmaxevent)
    ("status", "i", lambda d, N, nt, q, e: (N,)),
    ("llh", "d", lambda d, N, nt, q, e: (N,)),
    ("chi2", "d", lambda d, N, nt, q, e: (N,)),
    ("sllh", "d", lambda d, N, nt, q, e: (N, q)),
    ("x", "d", lambda d, N, nt, q, e: (N, nt, d.nx)),
    ("sx", "d", lambda d, N, nt, q, e: (N, nt, q, d.nx)),
    ("y", "d", lambda d, N, nt, q, e: (N, nt, d.ny)),
    ("sy", "d", lambda d, N, nt, q, e: (N, nt, q, d.ny)),
    ("sigmay", "d", lambda d, N, nt, q, e: (N, nt, d.ny)),
    ("ssigmay", "d", lambda d, N, nt, q, e: (N, nt, q, d.ny)),
    ("res", "d", lambda d, N, nt, q, e: (N, nt * d.nytrue)),
    ("sres", "d", lambda d, N, nt, q, e: (N, nt * d.nytrue, q)),
    ("FIM", "d", lambda d, N, nt, q, e: (N, q, q)),
    ("x0", "d", lambda d, N, nt, q, e: (N, d.nx)),
    ("sx0", "d", lambda d, N, nt, q, e: (N, q, d.nx)),
    ("xdot", "d", lambda d, N, nt, q, e: (N, d.nx_solver)),
    ("J", "d", lambda d, N, nt, q, e: (N, d.nx_solver, d.nx_solver)),
    ("z", "d", lambda d, N, nt, q, e: (N, e, d.nz)),
    ("sz", "d", lambda d, N, nt, q, e: (N, e, q, d.nz)),
    ("rz", "d", lambda d, N, nt, q, e: (N, e, d.nz)),
    ("srz", "d", lambda d, N, nt, q, e: (N, e, q, d.nz)),
    ("sigmaz", "d", lambda d, N, nt, q, e: (N, e, d.nz)),
    ("x_ss", "d", lambda d, N, nt, q, e: (N, d.nx)),
    ("sx_ss", "d", lambda d, N, nt, q, e: (N, q, d.nx)),
    ("preeq_status", "i", lambda d, N, nt, q, e: (N, 3)),
    ("preeq_numsteps", "i", lambda d, N, nt, q, e: (N, 3)),
    ("preeq_wrms", "d", lambda d, N, nt, q, e: (N,)),
    ("preeq_t", "d", lambda d, N, nt, q, e: (N,)),
    ("posteq_status", "i", lambda d, N, nt, q, e: (N, 3)),
    ("posteq_numsteps", "i", lambda d, N, nt, q, e: (N, 3)),
    ("posteq_wrms", "d", lambda d, N, nt, q, e: (N,)),
    ("posteq_t", "d", lambda d, N, nt, q, e: (N,)),
    ("numsteps", "i", lambda d, N, nt, q, e: (N, nt)),
    ("numrhsevals", "i", lambda d, N, nt, q, e: (N, nt)),
    ("numerrtestfails", "i", lambda d, N, nt, q, e: (N, nt)),
    ("numnonlinsolvconvfails", "i", lambda d, N, nt, q, e: (N, nt)),
    ("order", "i", lambda d, N, nt, q, e: (N, nt)),
    ("wall_seconds", "d", lambda d, N, nt, q, e: (1,)),
]


class RefBatch(C.Structure):
    _fields_ = [("N", C.c_int), ("nt", C.c_int), ("nplist", C.c_int),
                ("ts", _dp), ("plist", _ip), ("pscale", _ip), ("p", _dp), ("k", _dp),
                ("k_preeq", _dp), ("my", _dp), ("sigmay_in", _dp), ("mz", _dp),
                ("sigmaz_in", _dp), ("x0_in", _dp), ("sx0_in", _dp), ("k_presim", _dp), ("t_presim", _dp)] + [
        (n, _dp if t == "d" else _ip) for n, t, _ in _OUT_FIELDS]


def available(model: str) -> bool:
    return os.path.exists(os.path.join(_REF_DIR, f"libref_{model}.so"))


class RefModel:
    """The compiled reference for one pre-generated model."""

    def __init__(self, model: str):
        path = os.path.join(_REF_DIR, f"libref_{model}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing — run oracle/build_ref.sh")
        self.name = model
        self.lib = C.CDLL(path)
        self.lib.ref_run.restype = C.c_int
        self.lib.ref_max_threads.restype = C.c_int
        self.dims = RefDims()
        self.lib.ref_get_dims(C.byref(self.dims))
        d = self.dims
        p = np.zeros(d.np)
        k = np.zeros(max(d.nk, 1))
        self.lib.ref_get_nominal(p.ctypes.data_as(_dp), k.ctypes.data_as(_dp))
        self.p_nominal, self.k_nominal = p, k[: d.nk]

    def max_threads(self) -> int:
        """Host cores this process may use. Not omp_get_max_threads(): launchers such as torchrun export OMP_NUM_THREADS=1
        for their workers, which would time the reference on one core; run_simulations takes its thread count explicitly."""
        try:
            return max(1, len(os.sched_getaffinity(0)))
        except AttributeError:
            return max(1, os.cpu_count() or int(self.lib.ref_max_threads()))

    def run(self, ts, p, k, *, plist=None, pscale=None, k_preeq=None, my=None, sigmay=None, x0=None, sx0=None,
            k_presim=None, t_presim=None, reinit_all=False, reinit_states_sim=0, reinit_states_presim=0,
            mz=None, sigmaz=None, rtol=1e-8, atol=1e-12, rtol_fsa=np.nan, atol_fsa=np.nan,
            sensi_order=1, sensi_meth=1, ism=1, linsol=9, maxsteps=10000, newton_maxsteps=0,
            ss_mode=-1, ss_sensi_mode=-1, nmaxevent=-1, stldet=-1, reporting=0, num_threads=1,
            failfast=False, t0=0.0, ss_tol_factor=np.nan, rtol_ss=np.nan, atol_ss=np.nan,
            x_nonneg=0, maxtime=0.0, fields=None) -> dict:
        d = self.dims
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        N = p.shape[0]
        k = np.ascontiguousarray(np.atleast_2d(k), dtype=np.float64).reshape(N, d.nk)
        plist = np.arange(d.np, dtype=np.int32) if plist is None else np.ascontiguousarray(plist, dtype=np.int32)
        pscale = np.zeros(d.np, dtype=np.int32) if pscale is None else np.ascontiguousarray(pscale, dtype=np.int32)
        nt, q = len(ts), len(plist)
        ne = nmaxevent if nmaxevent >= 0 else 10
        opt = RefOptions(rtol=rtol, atol=atol, rtol_fsa=rtol_fsa, atol_fsa=atol_fsa,
                         ss_tol_factor=ss_tol_factor, rtol_ss=rtol_ss, atol_ss=atol_ss, t0=t0,
                         sensi_order=sensi_order, sensi_meth=sensi_meth, ism=ism, linsol=linsol,
                         maxsteps=maxsteps, newton_maxsteps=newton_maxsteps, ss_mode=ss_mode,
                         ss_sensi_mode=ss_sensi_mode, nmaxevent=nmaxevent, stldet=stldet,
                         reporting=reporting, num_threads=num_threads, failfast=int(failfast),
                         reinit_all=int(reinit_all), reinit_states_sim=int(reinit_states_sim),
                         reinit_states_presim=int(reinit_states_presim), x_nonneg=int(x_nonneg), maxtime=float(maxtime))
        b = RefBatch(N=N, nt=nt, nplist=q)
        keep = [ts, p, k, plist, pscale]
        b.ts, b.p, b.k = ts.ctypes.data_as(_dp), p.ctypes.data_as(_dp), k.ctypes.data_as(_dp)
        b.plist, b.pscale = plist.ctypes.data_as(_ip), pscale.ctypes.data_as(_ip)
        for name, arr, shape in (("k_preeq", k_preeq, (N, d.nk)), ("my", my, (N, nt, d.nytrue)),
                                 ("sigmay_in", sigmay, (N, nt, d.nytrue)),
                                 ("mz", mz, (N, ne, d.nztrue)), ("sigmaz_in", sigmaz, (N, ne, d.nztrue)),
                                 ("x0_in", x0, (N, d.nx)), ("sx0_in", sx0, (N, q, d.nx)),
                                 ("k_presim", k_presim, (N, d.nk)), ("t_presim", t_presim, (N,))):
            if arr is not None:
                a = np.ascontiguousarray(arr, dtype=np.float64).reshape(shape)
                keep.append(a)
                setattr(b, name, a.ctypes.data_as(_dp))
        out = {}
        for name, t, shp in _OUT_FIELDS:
            if fields is not None and name not in fields and name not in ("status", "wall_seconds"):
                continue
            shape = shp(d, N, nt, q, ne)
            a = np.zeros(shape, dtype=np.float64 if t == "d" else np.int32)
            out[name] = a
            setattr(b, name, a.ctypes.data_as(_dp if t == "d" else _ip))
        rc = self.lib.ref_run(C.byref(opt), C.byref(b))
        if rc != 0:
            raise RuntimeError(f"ref_run failed rc={rc}")
        out["wall_seconds"] = float(out["wall_seconds"][0])
        return out
