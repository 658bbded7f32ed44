"""amici_b200 — host-side mirror of ``amici.sim.sundials`` for the batched CUDA BDF engine.

The classes and functions here keep the names and argument meaning of the reference's Python
layer for the one path this package replaces:

* ``run_simulations(model, solver, edata_list, failfast=True, num_threads=1)`` and
  ``run_simulation`` (reference python/sdist/amici/sim/sundials/_swig_wrappers.py:46-133),
* ``Model`` / ``Solver`` / ``ExpData`` setters and getters that parametrise that call (reference
  include/amici/model.h, include/amici/solver.h:1869-2127, include/amici/edata.h),
* ``ReturnDataView`` field names and shapes (reference python/sdist/amici/sim/sundials/_numpy.py:414-459).

There is no CPU fallback: everything routes through the C ABI of ``lib/libamici_b200.so``
(include/amici_b200.h) and raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import enum
import math
import os
from typing import Sequence

import numpy as np

from . import _capi

__all__ = [
    "Model", "Solver", "ExpData", "ExpDataBatch", "ReturnDataView", "BatchResult", "run_simulation",
    "run_simulations", "runAmiciSimulations", "runAmiciSimulation", "SensitivityOrder", "SensitivityMethod",
    "InternalSensitivityMethod", "ParameterScaling", "RDataReporting", "LinearSolver", "SteadyStateComputationMode",
    "SteadyStateSensitivityMode", "SteadyStateStatus", "NewtonDampingFactorMode", "get_model",
    "available_models", "load_library", "AMICI_SUCCESS",
]

AMICI_SUCCESS = 0
AMICI_NOT_RUN = -1001
_LIB_PATH = os.environ.get("AMICI_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libamici_b200.so")
_lib = None


class SensitivityOrder(enum.IntEnum):
    none = 0
    first = 1
    second = 2


class SensitivityMethod(enum.IntEnum):
    none = 0
    forward = 1
    adjoint = 2


class InternalSensitivityMethod(enum.IntEnum):
    simultaneous = 1
    staggered = 2
    staggered1 = 3


class ParameterScaling(enum.IntEnum):
    none = 0
    ln = 1
    log10 = 2


class RDataReporting(enum.IntEnum):
    full = 0
    residuals = 1
    likelihood = 2
    observables_likelihood = 3


class LinearSolver(enum.IntEnum):
    dense = 1
    KLU = 9


class SteadyStateComputationMode(enum.IntEnum):
    newtonOnly = 0
    integrationOnly = 1
    integrateIfNewtonFails = 2


class SteadyStateSensitivityMode(enum.IntEnum):
    newtonOnly = 0
    integrationOnly = 1
    integrateIfNewtonFails = 2


class SteadyStateStatus(enum.IntEnum):
    failed_too_long_simulation = -5
    failed_damping = -4
    failed_factorization = -3
    failed_convergence = -2
    failed = -1
    not_run = 0
    success = 1


class NewtonDampingFactorMode(enum.IntEnum):
    off = 0
    on = 1


def load_library() -> C.CDLL:
    """Load the CUDA engine; fails loudly when it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("AMICI_B200_LIB", _LIB_PATH)   # override: experimental builds (tools/build_variant.py)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build the CUDA engine first (__graft_entry__.build()); "
                           "amici_b200 has no CPU fallback")
    lib = C.CDLL(path)
    lib.ab_model_count.restype = C.c_int
    lib.ab_model_name.restype = C.c_char_p
    lib.ab_model_name.argtypes = [C.c_int]
    lib.ab_get_dims.argtypes = [C.c_char_p, C.c_void_p]
    lib.ab_get_nominal.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    lib.ab_run_batch.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.ab_solve_device.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ab_reduce_objective.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ab_kernel_launch_count.restype = C.c_long
    lib.ab_last_error.restype = C.c_char_p
    lib.ab_last_kernel_name.restype = C.c_char_p
    _lib = lib
    return lib


class _Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "np", "nk", "ny", "nz", "ne", "nspl", "group_lanes", "cols_per_lane")]


def available_models() -> list[str]:
    lib = load_library()
    return [lib.ab_model_name(i).decode() for i in range(lib.ab_model_count())]


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(f"amici_b200: error {rc}: {load_library().ab_last_error().decode()}")


class Model:
    """Counterpart of ``amici.Model`` for a model compiled into the engine (settings container)."""

    def __init__(self, name: str):
        lib = load_library()
        d = _Dims()
        _check(lib.ab_get_dims(name.encode(), C.byref(d)))
        self.name = name
        self.nx_rdata = self.nx_solver = self.nx = d.nx
        self._np, self._nk = d.np, d.nk
        self.ny = self.nytrue = d.ny
        self.nz, self.ne, self.nspl = d.nz, d.ne, d.nspl
        self.group_lanes, self.cols_per_lane = d.group_lanes, d.cols_per_lane
        p = np.zeros(max(d.np, 1))
        k = np.zeros(max(d.nk, 1))
        _check(lib.ab_get_nominal(name.encode(), p.ctypes.data, k.ctypes.data))
        self._pscale = [ParameterScaling.none] * d.np
        self._p = list(p[: d.np])   # on parameter scale
        self._k = list(k[: d.nk])
        self._ts: list[float] = []
        self._plist = list(range(d.np))
        self._t0 = 0.0
        self._ss_mode = SteadyStateComputationMode.integrateIfNewtonFails
        self._ss_sensi_mode = SteadyStateSensitivityMode.integrateIfNewtonFails
        self._nmaxevent = 10
        self._reinit_all = False          # Model::set_reinitialize_fixed_parameter_initial_states
        self._reinit_idxs: list[int] = []  # Model::set_reinitialization_state_idxs

    # -- dimensions
    def np(self) -> int: return self._np
    def nk(self) -> int: return self._nk
    def nplist(self) -> int: return len(self._plist)
    def nt(self) -> int: return len(self._ts)
    def dims(self) -> dict: return dict(nx=self.nx, np=self._np, nk=self._nk, ny=self.ny)

    # -- settings (reference include/amici/model.h)
    def set_timepoints(self, ts: Sequence[float]):
        ts = [float(t) for t in ts]
        if any(b < a for a, b in zip(ts, ts[1:])):
            raise RuntimeError("Encountered non-monotonic timepoints, please order timepoints such that they are monotonically increasing!")
        self._ts = ts
    def get_timepoints(self): return list(self._ts)
    def set_t0(self, t0: float): self._t0 = float(t0)
    def t0(self): return self._t0
    def set_parameter_list(self, plist: Sequence[int]):
        plist = [int(i) for i in plist]
        if any(i < 0 or i >= self._np for i in plist):
            raise RuntimeError("Indices in plist must be in [0..np]")
        self._plist = plist
    def get_parameter_list(self): return list(self._plist)
    def set_parameter_scale(self, pscale):
        if isinstance(pscale, (int, enum.IntEnum)):
            pscale = [pscale] * self._np
        if len(pscale) != self._np:
            raise RuntimeError("Dimension mismatch. Size of parameter scaling does not match number of model parameters.")
        # keep unscaled values fixed, as Model::set_parameter_scale does
        un = self.get_unscaled_parameters()
        self._pscale = [ParameterScaling(int(s)) for s in pscale]
        self._p = [_scale(v, s) for v, s in zip(un, self._pscale)]
    def get_parameter_scale(self): return list(self._pscale)
    def set_free_parameters(self, p):
        if len(p) != self._np:
            raise RuntimeError("Dimension mismatch. Size of parameters does not match number of model parameters.")
        self._p = [float(v) for v in p]
    def get_free_parameters(self): return list(self._p)
    def get_unscaled_parameters(self): return [_unscale(v, s) for v, s in zip(self._p, self._pscale)]
    def set_fixed_parameters(self, k):
        if len(k) != self._nk:
            raise RuntimeError("Dimension mismatch. Size of fixedParameters does not match number of fixed model parameters.")
        self._k = [float(v) for v in k]
    def get_fixed_parameters(self): return list(self._k)
    # steady-state / event settings of amici::Model (reference include/amici/model.h)
    def set_steady_state_computation_mode(self, mode): self._ss_mode = SteadyStateComputationMode(int(mode))
    # re-initialisation of fixed-parameter-dependent initial states (reference src/model.cpp:1074-1097, :3110-3121)
    def set_reinitialize_fixed_parameter_initial_states(self, flag):
        self._reinit_all = bool(flag)
        if not flag:
            self._reinit_idxs = []
    def get_reinitialize_fixed_parameter_initial_states(self): return self._reinit_all or bool(self._reinit_idxs)
    def set_reinitialization_state_idxs(self, idxs):
        idxs = [int(i) for i in idxs]
        for i in idxs:
            if i < 0 or i >= self.nx:
                raise RuntimeError(f"Invalid state index given: {i}")
        self._reinit_idxs = idxs
    def get_reinitialization_state_idxs(self): return list(self._reinit_idxs)
    def get_steady_state_computation_mode(self): return self._ss_mode
    def set_steady_state_sensitivity_mode(self, mode): self._ss_sensi_mode = SteadyStateSensitivityMode(int(mode))
    def get_steady_state_sensitivity_mode(self): return self._ss_sensi_mode
    def set_state_is_non_negative(self, flags):
        flags = [bool(f) for f in flags]
        if len(flags) != self.nx:
            raise RuntimeError("Dimension of input stateIsNonNegative does not agree with number of state variables")
        if any(flags) and self.nx > 64:
            raise RuntimeError("amici_b200: state non-negativity is limited to models with nx <= 64")
        self._x_nonneg = sum(1 << i for i, f in enumerate(flags) if f)
    def get_state_is_non_negative(self): return [bool(getattr(self, "_x_nonneg", 0) >> i & 1) for i in range(self.nx)]
    def set_all_states_non_negative(self): self.set_state_is_non_negative([True] * self.nx)
    def set_n_max_event(self, n):
        if int(n) < 0:
            raise RuntimeError("nmaxevent must be non-negative")
        self._nmaxevent = int(n)
    def n_max_event(self): return self._nmaxevent
    def create_solver(self) -> "Solver": return Solver()

    def simulate(self, *, solver: "Solver | None" = None, edata=None, failfast: bool = True, num_threads: int = 1,
                 sensi_method=None, sensi_order=None):
        """``Model.simulate`` of the reference (python/sdist/amici/sim/sundials/_swig_wrappers.py:296-331): one ExpData (or
        none) -> one ReturnDataView, a sequence (or an ``ExpDataBatch``) -> the batch call; ``sensi_method`` /
        ``sensi_order`` (enum or its name) are applied to the solver for this call only."""
        solver = solver if solver is not None else self.create_solver()
        old_method = old_order = None
        if sensi_method is not None:
            old_method = solver.get_sensitivity_method()
            solver.set_sensitivity_method(SensitivityMethod[sensi_method] if isinstance(sensi_method, str) else sensi_method)
        if sensi_order is not None:
            old_order = solver.get_sensitivity_order()
            solver.set_sensitivity_order(SensitivityOrder[sensi_order] if isinstance(sensi_order, str) else sensi_order)
        try:
            if isinstance(edata, (list, tuple, ExpDataBatch)):
                return run_simulations(self, solver, edata, failfast=failfast, num_threads=num_threads)
            return run_simulation(self, solver, edata)
        finally:
            if old_method is not None:
                solver.set_sensitivity_method(old_method)
            if old_order is not None:
                solver.set_sensitivity_order(old_order)

    def get_name(self): return self.name
    def clone(self) -> "Model":
        import copy
        return copy.deepcopy(self)


def _unscale(v, s):
    return 10.0 ** v if s == ParameterScaling.log10 else math.exp(v) if s == ParameterScaling.ln else v


def _scale(v, s):
    return math.log10(v) if s == ParameterScaling.log10 else math.log(v) if s == ParameterScaling.ln else v


def get_model(name: str) -> Model:
    """Counterpart of the generated module's ``get_model()`` (reference models/*/wrapfunctions.cpp)."""
    return Model(name)


class Solver:
    """Settings of the batched BDF solver; setters/defaults follow ``amici::Solver`` (reference include/amici/solver.h:1869-2127)."""

    def __init__(self):
        self._rtol, self._atol = 1e-8, 1e-16
        self._rtol_fsa = self._atol_fsa = float("nan")
        self._sensi = SensitivityOrder.none
        self._sensi_meth = SensitivityMethod.forward
        self._ism = InternalSensitivityMethod.simultaneous
        self._maxsteps = 10000
        self._stldet = True
        self._max_nonlin_iters, self._max_conv_fails = 3, 10
        self._max_step_size = 0.0
        self._reporting = RDataReporting.full
        self._linsol = LinearSolver.KLU
        # steady-state solver (reference include/amici/solver.h:2040-2100)
        self._newton_maxsteps = 0
        self._newton_damping_mode = NewtonDampingFactorMode.on
        self._newton_damping_lb = 1e-8
        self._newton_step_ss_check = False
        self._sensi_ss_check = True
        self._ss_tol_factor = self._ss_tol_sensi_factor = 100.0
        self._rtol_ss = self._atol_ss = self._rtol_ss_sensi = self._atol_ss_sensi = float("nan")
        self._sensi_meth_preeq = SensitivityMethod.forward

    @staticmethod
    def _nonneg(v, what):
        if v < 0:
            raise RuntimeError(f"{what} must be a non-negative number")
        return float(v)

    def set_relative_tolerance(self, v): self._rtol = self._nonneg(v, "rtol")
    def get_relative_tolerance(self): return self._rtol
    def set_absolute_tolerance(self, v): self._atol = self._nonneg(v, "atol")
    def get_absolute_tolerance(self): return self._atol
    def set_relative_tolerance_fsa(self, v): self._rtol_fsa = self._nonneg(v, "rtol")
    def get_relative_tolerance_fsa(self): return self._rtol if math.isnan(self._rtol_fsa) else self._rtol_fsa
    def set_absolute_tolerance_fsa(self, v): self._atol_fsa = self._nonneg(v, "atol")
    def get_absolute_tolerance_fsa(self): return self._atol if math.isnan(self._atol_fsa) else self._atol_fsa
    def set_sensitivity_order(self, v): self._sensi = SensitivityOrder(int(v))
    def get_sensitivity_order(self): return self._sensi
    def set_sensitivity_method(self, v):
        # adjoint: event-free lane-group models, plain simulations (amici_b200/csrc/adjoint.cuh); the engine refuses the rest
        self._sensi_meth = SensitivityMethod(int(v))
    def get_sensitivity_method(self): return self._sensi_meth
    def set_internal_sensitivity_method(self, v):
        v = InternalSensitivityMethod(int(v))
        if v == InternalSensitivityMethod.staggered1:
            raise RuntimeError("amici_b200: InternalSensitivityMethod.staggered1 is not implemented (simultaneous and staggered are)")
        self._ism = v
    def get_internal_sensitivity_method(self): return self._ism
    def set_max_steps(self, v):
        if v <= 0:
            raise RuntimeError("maxsteps must be a positive number")
        self._maxsteps = int(v)
    def get_max_steps(self): return self._maxsteps
    def set_stability_limit_flag(self, v): self._stldet = bool(v)
    def get_stability_limit_flag(self): return self._stldet
    def set_max_nonlin_iters(self, v):
        if v < 0:
            raise RuntimeError("max_nonlin_iters must be non-negative")
        self._max_nonlin_iters = int(v)
    def get_max_nonlin_iters(self): return self._max_nonlin_iters
    def set_max_conv_fails(self, v):
        if v < 0:
            raise RuntimeError("max_conv_fails must be non-negative")
        self._max_conv_fails = int(v)
    def get_max_conv_fails(self): return self._max_conv_fails
    def set_max_step_size(self, v):
        if v < 0:
            raise RuntimeError("max_step_size must be non-negative.")
        self._max_step_size = float(v)
    def get_max_step_size(self): return self._max_step_size
    def set_return_data_reporting_mode(self, v):
        self._reporting = RDataReporting(int(v))
    def get_return_data_reporting_mode(self): return self._reporting
    def set_newton_max_steps(self, v):
        if v < 0:
            raise RuntimeError("newton_maxsteps must be a non-negative number")
        self._newton_maxsteps = int(v)
    def get_newton_max_steps(self): return self._newton_maxsteps
    def set_newton_damping_factor_mode(self, v): self._newton_damping_mode = NewtonDampingFactorMode(int(v))
    def get_newton_damping_factor_mode(self): return self._newton_damping_mode
    def set_newton_damping_factor_lower_bound(self, v): self._newton_damping_lb = float(v)
    def get_newton_damping_factor_lower_bound(self): return self._newton_damping_lb
    def set_newton_step_steady_state_check(self, flag):
        if flag:
            raise RuntimeError("amici_b200: the Newton-step steady-state convergence check is not implemented")
        self._newton_step_ss_check = False
    def get_newton_step_steady_state_check(self): return self._newton_step_ss_check
    def set_sensi_steady_state_check(self, flag): self._sensi_ss_check = bool(flag)
    def get_sensi_steady_state_check(self): return self._sensi_ss_check
    def set_steady_state_tolerance_factor(self, v): self._ss_tol_factor = self._nonneg(v, "ss_tol_factor")
    def get_steady_state_tolerance_factor(self): return self._ss_tol_factor
    def set_steady_state_sensi_tolerance_factor(self, v): self._ss_tol_sensi_factor = self._nonneg(v, "ss_tol_sensi_factor")
    def get_steady_state_sensi_tolerance_factor(self): return self._ss_tol_sensi_factor
    def set_relative_tolerance_steady_state(self, v): self._rtol_ss = self._nonneg(v, "rtol")
    def get_relative_tolerance_steady_state(self):
        return self._rtol * self._ss_tol_factor if math.isnan(self._rtol_ss) else self._rtol_ss
    def set_absolute_tolerance_steady_state(self, v): self._atol_ss = self._nonneg(v, "atol")
    def get_absolute_tolerance_steady_state(self):
        return self._atol * self._ss_tol_factor if math.isnan(self._atol_ss) else self._atol_ss
    def set_relative_tolerance_steady_state_sensi(self, v): self._rtol_ss_sensi = self._nonneg(v, "rtol")
    def get_relative_tolerance_steady_state_sensi(self):
        return self._rtol * self._ss_tol_sensi_factor if math.isnan(self._rtol_ss_sensi) else self._rtol_ss_sensi
    def set_absolute_tolerance_steady_state_sensi(self, v): self._atol_ss_sensi = self._nonneg(v, "atol")
    def get_absolute_tolerance_steady_state_sensi(self):
        return self._atol * self._ss_tol_sensi_factor if math.isnan(self._atol_ss_sensi) else self._atol_ss_sensi
    def set_sensitivity_method_pre_equilibration(self, v):
        v = SensitivityMethod(int(v))
        if v == SensitivityMethod.adjoint:
            raise RuntimeError("amici_b200: adjoint pre-equilibration is outside the batched forward path")
        self._sensi_meth_preeq = v
    def get_sensitivity_method_pre_equilibration(self): return self._sensi_meth_preeq
    def set_max_time(self, seconds):
        if seconds < 0:
            raise RuntimeError("maxtime must be non-negative")
        self._maxtime = float(seconds)
    def get_max_time(self): return getattr(self, "_maxtime", 0.0)
    def set_linear_solver(self, v): self._linsol = LinearSolver(int(v))  # accepted; the engine always uses its batched LU
    def get_linear_solver(self): return self._linsol
    def clone(self):
        import copy
        return copy.deepcopy(self)

    def _options(self, t0: float, model: "Model | None" = None) -> _capi.SolverOptions:
        fsa = self._sensi >= SensitivityOrder.first and self._sensi_meth in (SensitivityMethod.forward, SensitivityMethod.adjoint)
        if self._sensi >= SensitivityOrder.second:
            raise RuntimeError("amici_b200: second-order sensitivities are not implemented")
        return _capi.default_options(
            rtol=self._rtol, atol=self._atol, rtol_fsa=self.get_relative_tolerance_fsa(),
            atol_fsa=self.get_absolute_tolerance_fsa(), t0=t0,
            hmax_inv=(1.0 / self._max_step_size if self._max_step_size > 0 else 0.0), sensi=1 if fsa else 0,
            ism=int(self._ism), maxsteps=self._maxsteps, stldet=int(self._stldet),
            max_nonlin_iters=self._max_nonlin_iters, max_conv_fails=self._max_conv_fails, max_err_test_fails=7,
            reporting=int(self._reporting),
            rtol_ss=self.get_relative_tolerance_steady_state(), atol_ss=self.get_absolute_tolerance_steady_state(),
            rtol_ss_sensi=self.get_relative_tolerance_steady_state_sensi(),
            atol_ss_sensi=self.get_absolute_tolerance_steady_state_sensi(),
            newton_damping_lower_bound=self._newton_damping_lb, newton_maxsteps=self._newton_maxsteps,
            newton_damping_mode=int(self._newton_damping_mode), sensi_ss_check=int(self._sensi_ss_check),
            ss_mode=int(model._ss_mode) if model is not None else 2,
            ss_sensi_mode=int(model._ss_sensi_mode) if model is not None else 2,
            nmaxevent=model._nmaxevent if model is not None else 10,
            x_nonneg=getattr(model, "_x_nonneg", 0) if model is not None else 0, maxtime=self.get_max_time(),
            sensi_meth=int(self._sensi_meth) if fsa else 0)


class ExpData:
    """One experimental condition (``amici::ExpData``, reference include/amici/edata.h:20)."""

    def __init__(self, model: Model):
        self._ny = model.ny
        self._nz, self._nmaxevent = model.nz, model.n_max_event()
        self._mz = np.full(self._nmaxevent * self._nz, np.nan)
        self._sigmaz = np.full(self._nmaxevent * self._nz, np.nan)
        self.ts_ = model.get_timepoints()
        self.free_parameters: list[float] = []   # empty = use the model's
        self.fixed_parameters: list[float] = []
        self.fixed_parameters_pre_equilibration: list[float] = []   # non-empty = pre-equilibrate (edata.h)
        self.fixed_parameters_presimulation: list[float] = []       # k during the presimulation; empty = the simulation's
        self.t_presim: float = 0.0                                   # > 0: integrate [t0 - t_presim, t0] first (edata.h)
        # ExpData(Model const&) (reference src/edata.cpp:60-72)
        m_idxs = list(getattr(model, "_reinit_idxs", []))
        self.reinitialize_fixed_parameter_initial_states: bool = bool(getattr(model, "_reinit_all", False)) and not m_idxs
        self.reinitialization_state_idxs_sim: list[int] = m_idxs
        self.reinitialization_state_idxs_presim: list[int] = []
        self.pscale: list = []
        self.plist: list[int] = []
        self.x0: list[float] = []    # initial state override (edata.h); empty = the model's x0(p, k)
        self.sx0: list[float] = []   # [nplist][nx] initial sensitivities on the parameter scale; empty = the model's
        n = len(self.ts_) * self._ny
        self._my = np.full(n, np.nan)
        self._sigmay = np.full(n, np.nan)

    @classmethod
    def from_return_data(cls, model: Model, rdata, sigma_y, sigma_z=(), seed: int = -1) -> "ExpData":
        """Synthetic data: measurements = simulated observables + N(0, sigma^2) (``ExpData(rdata, sigma_y, sigma_z, seed)``,
        reference src/edata.cpp:83-131).  ``sigma_y`` has ny or nt*ny entries, ``sigma_z`` nz or nmaxevent*nz; a negative
        seed draws a fresh one.  The noise comes from numpy's generator, not std::mt19937: same distribution, other stream."""
        e = cls(model)
        e.set_timepoints(list(np.asarray(rdata["ts"], dtype=float)))
        nt, ny, nz, nme = len(e.ts_), e._ny, e._nz, e._nmaxevent
        sy, sz = np.asarray(sigma_y, dtype=float).ravel(), np.asarray(sigma_z, dtype=float).ravel()
        if sy.size not in (ny, ny * nt):
            raise RuntimeError(f"Dimension of sigma_y must be {ny} or {ny * nt}, was {sy.size}")
        if sz.size not in (nz, nz * nme):
            raise RuntimeError(f"Dimension of sigma_z must be {nz} or {nz * nme}, was {sz.size}")
        if (sy.size and not (sy > 0).all()) or (sz.size and not (sz > 0).all()):
            raise RuntimeError("Encountered sigma <= 0")
        rng = np.random.default_rng(None if seed < 0 else seed)
        sig_y = np.broadcast_to(sy.reshape(-1, ny), (nt, ny))
        e.set_measurements(np.asarray(rdata["y"], dtype=float).reshape(nt, ny) + sig_y * rng.standard_normal((nt, ny)))
        e.set_noise_scales(sig_y)
        if nz:
            sig_z = np.broadcast_to(sz.reshape(-1, nz), (nme, nz))
            e.set_event_measurements(np.asarray(rdata["z"], dtype=float).reshape(nme, nz) + sig_z * rng.standard_normal((nme, nz)))
            e.set_event_noise_scales(sig_z)
        return e

    def nt(self): return len(self.ts_)
    def set_timepoints(self, ts):
        self.ts_ = [float(t) for t in ts]
        n = len(self.ts_) * self._ny
        self._my = np.full(n, np.nan)
        self._sigmay = np.full(n, np.nan)
    def get_timepoints(self): return list(self.ts_)
    def set_measurements(self, my):
        my = np.asarray(my, dtype=np.float64).ravel()
        if my.size != self.nt() * self._ny:
            raise RuntimeError(f"Input observedData did not match dimensions nt ({self.nt()}) x nytrue ({self._ny}), was {my.size}.")
        self._my = my.copy()
    def get_measurements(self): return self._my.copy()
    def set_noise_scales(self, s):
        s = np.asarray(s, dtype=np.float64).ravel()
        if s.size == 1:
            s = np.full(self.nt() * self._ny, float(s[0]))
        if s.size != self.nt() * self._ny:
            raise RuntimeError(f"Input observedDataStdDev did not match dimensions nt ({self.nt()}) x nytrue ({self._ny}), was {s.size}.")
        if np.any(s[~np.isnan(s)] <= 0):
            raise RuntimeError("Encountered sigma <= 0 in sigmay! value: %g" % float(s[~np.isnan(s)].min()))
        self._sigmay = s.copy()
    def get_noise_scales(self): return self._sigmay.copy()
    def set_event_measurements(self, mz):
        mz = np.asarray(mz, dtype=np.float64).ravel()
        if mz.size != self._nmaxevent * self._nz:
            raise RuntimeError(f"Input observedEvents did not match dimensions nmaxevent ({self._nmaxevent}) x nztrue ({self._nz}), was {mz.size}.")
        self._mz = mz.copy()
    def get_event_measurements(self): return self._mz.copy()
    def set_event_noise_scales(self, s):
        s = np.asarray(s, dtype=np.float64).ravel()
        if s.size == 1:
            s = np.full(self._nmaxevent * self._nz, float(s[0]))
        if s.size != self._nmaxevent * self._nz:
            raise RuntimeError(f"Input observedEventsStdDev did not match dimensions nmaxevent ({self._nmaxevent}) x nztrue ({self._nz}), was {s.size}.")
        if np.any(s[~np.isnan(s)] <= 0):
            raise RuntimeError("Encountered sigma <= 0 in sigmaz! value: %g" % float(s[~np.isnan(s)].min()))
        self._sigmaz = s.copy()
    def get_event_noise_scales(self): return self._sigmaz.copy()


def _reinit_key(e) -> tuple:
    """(simulation, presimulation) re-initialisation selection of one ExpData: "all" or a tuple of state indices
    (ConditionContext, reference src/edata.cpp:492-548)."""
    if e.reinitialize_fixed_parameter_initial_states:
        return ("all", "all")
    return (tuple(int(i) for i in e.reinitialization_state_idxs_sim), tuple(int(i) for i in e.reinitialization_state_idxs_presim))


def _reinit_mask(sel, model, presim: bool = False) -> int:
    if sel is None:   # the model's settings: its index list applies to the simulation context only
        sel = True if (model._reinit_all and not model._reinit_idxs) else ([] if presim else model._reinit_idxs)
    if sel is True:
        return (1 << 64) - 1
    mask = 0
    for i in sel:
        if int(i) < 0 or int(i) >= model.nx:
            raise RuntimeError(f"Invalid state index given: {int(i)}")
        if int(i) >= 64:
            raise RuntimeError("amici_b200: state re-initialisation is limited to models with nx <= 64")
        mask |= 1 << int(i)
    return mask


class ExpDataBatch:
    """Struct-of-arrays form of ``list[ExpData]`` for conditions that share ts / plist / pscale.

    This is what the C ABI consumes (``ab_batch``); ``run_simulations`` builds it from a list of
    ``ExpData`` or accepts it directly, which avoids one Python object per condition for large batches.
    """

    def __init__(self, ts, p, k, my=None, sigmay=None, plist=None, pscale=None, k_preeq=None, mz=None, sigmaz=None,
                 x0=None, sx0=None, k_presim=None, t_presim=None,
                 reinit_states_sim=None, reinit_states_presim=None):
        self.ts = np.ascontiguousarray(ts, dtype=np.float64)
        self.p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        self.k = np.ascontiguousarray(np.atleast_2d(k), dtype=np.float64)
        self.my = None if my is None else np.ascontiguousarray(my, dtype=np.float64)
        self.sigmay = None if sigmay is None else np.ascontiguousarray(sigmay, dtype=np.float64)
        self.plist = None if plist is None else np.ascontiguousarray(plist, dtype=np.int32)
        self.pscale = None if pscale is None else np.ascontiguousarray(pscale, dtype=np.int32)
        self.k_preeq = None if k_preeq is None else np.ascontiguousarray(np.atleast_2d(k_preeq), dtype=np.float64)
        self.mz = None if mz is None else np.ascontiguousarray(mz, dtype=np.float64)
        self.sigmaz = None if sigmaz is None else np.ascontiguousarray(sigmaz, dtype=np.float64)
        # ExpData::x0 / sx0: initial state (and its sensitivities, on the parameter scale) instead of the model's
        self.x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        self.sx0 = None if sx0 is None else np.ascontiguousarray(sx0, dtype=np.float64)
        # state re-initialisation after pre-equilibration / presimulation: index lists, True = all states, None = the model's
        self.reinit_states_sim, self.reinit_states_presim = reinit_states_sim, reinit_states_presim
        # ExpData::t_presim / fixed_parameters_presimulation: per condition, 0 = no presimulation
        self.t_presim = None if t_presim is None else np.ascontiguousarray(np.broadcast_to(np.asarray(t_presim, dtype=np.float64), (self.p.shape[0],)))
        self.k_presim = None if k_presim is None else np.ascontiguousarray(np.atleast_2d(k_presim), dtype=np.float64)
        if self.k_presim is not None and self.t_presim is None:
            raise RuntimeError("amici_b200: k_presim without t_presim")

    def __len__(self): return self.p.shape[0]

    @staticmethod
    def from_list(model: Model, edatas: Sequence[ExpData | None]) -> "ExpDataBatch":
        N = len(edatas)
        ts = None
        plist = pscale = None
        p = np.empty((N, model.np()))
        k = np.empty((N, model.nk()))
        for i, e in enumerate(edatas):
            e_ts = model.get_timepoints() if e is None or not e.ts_ else e.ts_
            e_plist = model.get_parameter_list() if e is None or not e.plist else list(e.plist)
            e_pscale = [int(s) for s in (model.get_parameter_scale() if e is None or not e.pscale else e.pscale)]
            if ts is None:
                ts, plist, pscale = e_ts, e_plist, e_pscale
            elif list(e_ts) != list(ts) or e_plist != plist or e_pscale != pscale:
                raise RuntimeError("amici_b200: all conditions of one batch must share timepoints, plist and pscale")
            p[i] = model.get_free_parameters() if e is None or not len(e.free_parameters) else e.free_parameters
            k[i] = model.get_fixed_parameters() if e is None or not len(e.fixed_parameters) else e.fixed_parameters
        nt = len(ts)
        my = np.full((N, nt, model.ny), np.nan)
        sg = np.full((N, nt, model.ny), np.nan)
        for i, e in enumerate(edatas):
            if e is not None:
                my[i] = e._my.reshape(nt, model.ny)
                sg[i] = e._sigmay.reshape(nt, model.ny)
        has_preeq = [e is not None and len(e.fixed_parameters_pre_equilibration) > 0 for e in edatas]
        k_preeq = None
        if any(has_preeq):
            if not all(has_preeq):
                raise RuntimeError("amici_b200: either all or none of the conditions of one batch pre-equilibrate")
            k_preeq = np.array([e.fixed_parameters_pre_equilibration for e in edatas], dtype=np.float64)
            if k_preeq.shape != (N, model.nk()):
                raise RuntimeError("Dimension mismatch. Size of fixed_parameters_pre_equilibration does not match nk.")
        mz = sz = None
        if model.nz > 0 and any(e is not None and not np.isnan(e._mz).all() for e in edatas):
            nme = model.n_max_event()
            mz = np.full((N, nme, model.nz), np.nan)
            sz = np.full((N, nme, model.nz), np.nan)
            for i, e in enumerate(edatas):
                if e is not None:
                    if e._nmaxevent != nme:
                        raise RuntimeError("amici_b200: ExpData was created with a different n_max_event than the model has now")
                    mz[i] = e._mz.reshape(nme, model.nz)
                    sz[i] = e._sigmaz.reshape(nme, model.nz)
        x0 = sx0 = None
        has_x0 = [e is not None and len(e.x0) > 0 for e in edatas]
        has_sx0 = [e is not None and len(e.sx0) > 0 for e in edatas]
        if any(has_x0) != all(has_x0) or any(has_sx0) != all(has_sx0):
            raise RuntimeError("amici_b200: either all or none of the conditions of one batch give x0 / sx0")
        if any(has_x0):
            x0 = np.array([e.x0 for e in edatas], dtype=np.float64)
            if x0.shape != (N, model.nx):
                raise RuntimeError("Dimension mismatch. Size of x0 does not match number of model states.")
        if any(has_sx0):
            sx0 = np.array([np.asarray(e.sx0, dtype=np.float64).reshape(-1) for e in edatas])
            if sx0.shape != (N, len(plist) * model.nx):
                raise RuntimeError("Dimension mismatch. Size of sx0 does not match number of model states * number of "
                                   "parameter selected for sensitivities.")
            sx0 = sx0.reshape(N, len(plist), model.nx)
        t_presim = k_presim = None
        if any(e is not None and e.t_presim > 0 for e in edatas):
            t_presim = np.array([0.0 if e is None else float(e.t_presim) for e in edatas])
            k_presim = k.copy()   # ConditionContext(presimulation) without dedicated parameters keeps the simulation's
            for i, e in enumerate(edatas):
                if e is not None and len(e.fixed_parameters_presimulation):
                    if len(e.fixed_parameters_presimulation) != model.nk():
                        raise RuntimeError("Number of fixed parameters in model does not match ExpData (presimulation).")
                    k_presim[i] = e.fixed_parameters_presimulation
        reinit = {_reinit_key(e) for e in edatas if e is not None}
        if len(reinit) > 1:
            raise RuntimeError("amici_b200: all conditions of one batch must share the state re-initialisation settings")
        r_sim, r_presim = reinit.pop() if reinit else (None, None)
        return ExpDataBatch(ts, p, k, my, sg, plist, pscale, k_preeq=k_preeq, mz=mz, sigmaz=sz, x0=x0, sx0=sx0,
                            k_presim=k_presim, t_presim=t_presim,
                            reinit_states_sim=True if r_sim == "all" else None if r_sim is None else list(r_sim),
                            reinit_states_presim=True if r_presim == "all" else None if r_presim is None else list(r_presim))


_RDATA_ALIASES = {"num_rhs_evals": "numrhsevals", "num_err_test_fails": "numerrtestfails",
                  "num_non_lin_solv_conv_fails": "numnonlinsolvconvfails", "t": "ts"}


class BatchResult:
    """All ReturnData fields of a batch as arrays with the condition index first."""

    def __init__(self, fields: dict, ts, nplist: int, kernel_ms: float | None = None):
        self.fields = fields
        self.ts = np.asarray(ts)
        self.nplist = nplist
        self.kernel_ms = kernel_ms

    def __getitem__(self, key): return self.fields[_RDATA_ALIASES.get(key, key)]
    def __contains__(self, key): return _RDATA_ALIASES.get(key, key) in self.fields
    def __len__(self): return len(self.fields["status"])

    def release(self):
        """Return the page-locked output buffers to the pool; the arrays of this result must not be used afterwards."""
        keep, self._keep = getattr(self, "_keep", []), []
        self.fields = {}
        _recycle(keep)
    def view(self, i: int) -> "ReturnDataView": return ReturnDataView(self, i)
    def views(self) -> list["ReturnDataView"]: return [ReturnDataView(self, i) for i in range(len(self))]


class ReturnDataView:
    """Per-condition view with the reference's field names/shapes (reference _numpy.py:414-459)."""

    def __init__(self, batch: BatchResult, i: int):
        self._b, self._i = batch, i

    def __getitem__(self, key):
        key = _RDATA_ALIASES.get(key, key)
        if key == "ts":
            return self._b.ts
        v = self._b.fields[key][self._i]
        return v.item() if np.ndim(v) == 0 else v

    def __getattr__(self, key):
        if key.startswith("_"):
            raise AttributeError(key)
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def keys(self): return ["ts"] + list(self._b.fields.keys())
    def __contains__(self, key): return key == "ts" or key in self._b


_PINNED_POOL: dict = {}   # (shape, dtype) -> [torch tensors]: page-locked output buffers are recycled explicitly


def _pinned_empty(shape, dtype):
    """Host output buffer; page-locked through torch when available so D2H copies run at full PCIe rate and
    overlap the solve. Buffers return to a pool when their BatchResult is garbage-collected (pinning 2.5 GB
    costs seconds, so it must not happen per call)."""
    try:
        import torch
        if torch.cuda.is_available():
            key = (tuple(int(s) for s in shape), np.dtype(dtype).str)
            pool = _PINNED_POOL.get(key)
            if pool:
                t = pool.pop()
            else:
                tdt = torch.float64 if dtype == np.float64 else torch.int32
                t = torch.empty(key[0], dtype=tdt, pin_memory=True)
            return t.numpy(), (key, t)
    except Exception:  # pragma: no cover - pinned memory is an optimisation only
        pass
    return np.empty(shape, dtype=dtype), None


def _recycle(keep):
    for item in keep:
        if isinstance(item, tuple) and len(item) == 2 and isinstance(item[0], tuple):
            _PINNED_POOL.setdefault(item[0], []).append(item[1])


def run_batch(model: Model, solver: Solver, batch: ExpDataBatch, fields: Sequence[str] | None = None) -> BatchResult:
    """Run one ``ab_run_batch`` call on host arrays (the drop-in for the reference's batch call)."""
    lib = load_library()
    N, nt = len(batch), len(batch.ts)
    plist = np.asarray(model.get_parameter_list(), dtype=np.int32) if batch.plist is None else batch.plist
    pscale = np.asarray([int(s) for s in model.get_parameter_scale()], dtype=np.int32) if batch.pscale is None else batch.pscale
    opts = solver._options(model.t0(), model)
    # t = inf output points: post-equilibration after the last finite timepoint (reference src/forwardproblem.cpp:376-387)
    finite = np.isfinite(batch.ts)
    has_posteq = bool(len(batch.ts)) and not finite.all()
    if has_posteq and (np.isnan(batch.ts).any() or (batch.ts == -np.inf).any() or finite[int(np.argmin(finite)):].any()):
        raise RuntimeError("amici_b200: timepoints must be finite, followed by t = inf entries only")
    opts.posteq = int(has_posteq)
    opts.reinit_states_sim = _reinit_mask(batch.reinit_states_sim, model)
    opts.reinit_states_presim = _reinit_mask(batch.reinit_states_presim, model, presim=True)
    q = len(plist) if opts.sensi else 0
    d = dict(model.dims(), nz=model.nz, nmaxevent=opts.nmaxevent)
    b = _capi.BatchView(N=N, nt=nt, nplist=len(plist))
    keep = [batch.ts, batch.p, batch.k, plist, pscale]
    if batch.p.shape != (N, d["np"]) or batch.k.shape != (N, d["nk"]):
        raise RuntimeError("amici_b200: parameter array shapes do not match the model")
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    b.ts, b.p, b.k = batch.ts.ctypes.data_as(dp), batch.p.ctypes.data_as(dp), batch.k.ctypes.data_as(dp)
    b.plist, b.pscale = plist.ctypes.data_as(ip), pscale.ctypes.data_as(ip)
    if batch.my is not None:
        b.my = batch.my.ctypes.data_as(dp)
    if batch.sigmay is not None:
        b.sigmay_in = batch.sigmay.ctypes.data_as(dp)
    if batch.x0 is not None:
        if batch.x0.shape != (N, d["nx"]):
            raise RuntimeError("Dimension mismatch. Size of x0 does not match number of model states.")
        b.x0_in = batch.x0.ctypes.data_as(dp)
    if batch.sx0 is not None and opts.sensi:
        if batch.sx0.shape != (N, len(plist), d["nx"]):
            raise RuntimeError("Dimension mismatch. Size of sx0 does not match number of model states * number of "
                               "parameter selected for sensitivities.")
        b.sx0_in = batch.sx0.ctypes.data_as(dp)
    if batch.k_preeq is not None:
        if batch.k_preeq.shape != (N, d["nk"]):
            raise RuntimeError("amici_b200: k_preeq shape does not match the model")
        b.k_preeq = batch.k_preeq.ctypes.data_as(dp)
    if batch.t_presim is not None and (batch.t_presim > 0).any():
        b.t_presim = batch.t_presim.ctypes.data_as(dp)
        if batch.k_presim is not None:
            if batch.k_presim.shape != (N, d["nk"]):
                raise RuntimeError("Number of fixed parameters in model does not match ExpData (presimulation).")
            b.k_presim = batch.k_presim.ctypes.data_as(dp)
    for name, arr in (("mz", batch.mz), ("sigmaz_in", batch.sigmaz)):
        if arr is not None:
            if arr.shape != (N, opts.nmaxevent, model.nz):
                raise RuntimeError(f"amici_b200: {name} must have shape [N][nmaxevent][nz]")
            setattr(b, name, arr.ctypes.data_as(dp))
    # which ReturnData fields a reporting mode holds (ReturnData::initialize_*_reporting, reference src/rdata.cpp:104-211); the
    # per-timepoint solver counters are kept in every mode (the reference has them in `full` only; bench.py reads them)
    counters = {"status", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails", "order"}
    mode_fields = {
        int(RDataReporting.likelihood): counters | {"llh", "sllh", "FIM"},
        int(RDataReporting.observables_likelihood): counters | {"llh", "sllh", "FIM", "y", "sigmay", "sy", "ssigmay"},
        int(RDataReporting.residuals): counters | {"llh", "chi2", "y", "sigmay", "sy", "ssigmay", "res", "sres"},
    }.get(opts.reporting)
    likelihood_only = mode_fields is not None
    lk_fields = mode_fields or set()
    out = {}
    for name, t, shp in _capi.OUT_FIELDS:
        if fields is not None and name not in fields and name != "status":
            continue
        if likelihood_only and name not in lk_fields:
            continue
        if name in _capi.EVENT_FIELDS and model.nz == 0:
            continue
        if name in _capi.POSTEQ_FIELDS and not has_posteq:
            # ReturnData defaults without post-equilibration (reference src/rdata.cpp:172-173)
            shape = shp(d, N, nt, q)
            out[name] = (np.zeros(shape, dtype=np.int32) if t == "i" else np.full(shape, np.nan))
            continue
        if name in _capi.PREEQ_FIELDS and batch.k_preeq is None:
            # ReturnData defaults without pre-equilibration (reference src/rdata.cpp:170-171, :192-196)
            shape = shp(d, N, nt, q)
            out[name] = (np.zeros(shape, dtype=np.int32) if t == "i" else np.full(shape, np.nan))
            continue
        shape = shp(d, N, nt, q)
        if q == 0 and name in ("sllh", "sx", "sy", "ssigmay", "sres", "FIM", "sx0", "sx_ss", "sz", "srz"):
            out[name] = np.zeros(shape)
            continue
        if opts.sensi_meth == 2 and name in ("sx", "sy", "ssigmay", "sres", "FIM", "sx0", "sx_ss", "sz", "srz"):
            # adjoint sensitivity analysis gives the gradient sllh only (the reference leaves these fields empty)
            out[name] = np.full(shape, np.nan)
            continue
        a, t_keep = _pinned_empty(shape, np.float64 if t == "d" else np.int32)
        keep.append(t_keep)
        out[name] = a
        setattr(b, name, a.ctypes.data_as(dp if t == "d" else ip))
    kms = C.c_float(0.0)
    _check(lib.ab_run_batch(model.name.encode(), C.byref(opts), C.byref(b), C.byref(kms)))
    if likelihood_only and (fields is None or "chi2" in fields or "llh" in fields):
        # the objective scalars a mode does not compute keep the reference's defaults (src/rdata.cpp:104-107, :970-982)
        if "chi2" not in lk_fields:
            out["chi2"] = np.full(N, np.nan)
    res = BatchResult(out, batch.ts, q, float(kms.value))
    res._keep = keep
    return res


def run_simulations(model: Model, solver: Solver, edata_list, failfast: bool = True, num_threads: int = 1):
    """Drop-in for ``amici.sim.sundials.run_simulations`` (reference _swig_wrappers.py:87-133).

    ``edata_list`` is a list of ``ExpData`` (returns ``list[ReturnDataView]`` like the reference) or an
    ``ExpDataBatch`` (returns a ``BatchResult``).  ``num_threads`` is accepted for signature
    compatibility; the batch runs on the current CUDA device.  ``failfast``: the whole batch is
    integrated concurrently, so — as in the reference with several threads (src/amici.cpp:246-260) —
    which conditions are skipped after a failure is not deterministic there; here none are skipped.
    """
    del failfast, num_threads
    if isinstance(edata_list, ExpDataBatch):
        return run_batch(model, solver, edata_list)
    if len(edata_list) == 0:
        return []
    # The reference lets every ExpData carry its own timepoints, parameter list, scales and pre-equilibration condition
    # (ConditionContext, reference src/edata.cpp:438-556).  Conditions that agree on those are integrated as one batch;
    # the results go back in the caller's order.
    out: list = [None] * len(edata_list)
    for idx, batch in group_edatas(model, edata_list):
        views = run_batch(model, solver, batch).views()
        for i, v in zip(idx, views):
            out[i] = v
    return out


def group_edatas(model: Model, edatas: Sequence[ExpData | None]) -> list[tuple[list[int], ExpDataBatch]]:
    """Partition a list of ExpData into batches of conditions that share (timepoints, plist, pscale, pre-equilibration
    yes/no) — what one ``ab_run_batch`` call requires. Returns ``[(indices into edatas, ExpDataBatch), ...]`` in order of
    first appearance."""
    groups: dict = {}
    for i, e in enumerate(edatas):
        ts = tuple(model.get_timepoints() if e is None or not e.ts_ else e.ts_)
        plist = tuple(model.get_parameter_list() if e is None or not e.plist else e.plist)
        pscale = tuple(int(s) for s in (model.get_parameter_scale() if e is None or not e.pscale else e.pscale))
        preeq = e is not None and len(e.fixed_parameters_pre_equilibration) > 0
        init = (e is not None and len(e.x0) > 0, e is not None and len(e.sx0) > 0)
        reinit = None if e is None else _reinit_key(e)
        groups.setdefault((ts, plist, pscale, preeq, init, reinit), []).append(i)
    return [(idx, ExpDataBatch.from_list(model, [edatas[i] for i in idx])) for idx in groups.values()]


def run_simulation(model: Model, solver: Solver, edata: ExpData | None = None) -> ReturnDataView:
    """Drop-in for ``amici.sim.sundials.run_simulation`` (reference _swig_wrappers.py:46-84)."""
    return run_batch(model, solver, ExpDataBatch.from_list(model, [edata])).view(0)


# names used by older AMICI releases (BASELINE.json quotes them)
runAmiciSimulations = run_simulations
runAmiciSimulation = run_simulation
