"""Sharded objective: llh and its gradient w.r.t. *problem* parameters for a set of conditions with a parameter mapping.

This is the consumer of the batched solve that the reference implements in Python around ``run_simulations``:
``simulate_petab`` fills every condition's parameter vector from the problem parameters through the PEtab parameter mapping
(reference python/sdist/amici/sim/sundials/petab/v1/_simulations.py:201-219), sums ``llh`` over conditions (:221-222) and
accumulates each condition's ``sllh`` entries onto the problem parameter they are mapped from (``aggregate_sllh``, :264-359).
The PEtab library itself is not needed for that arithmetic, so the mapping is taken in array form:

* ``index[n, i] >= 0``: model parameter ``i`` of condition ``n`` is problem parameter ``index[n, i]``
  (``map_sim_var[model_par] = problem_par``);
* ``index[n, i] == -1``: it is the numeric value ``value[n, i]`` (a condition-table override), and carries no gradient.

Everything stays on the device: the fill is a gather, the aggregation an ``index_add_``, and with several ranks each one
holds a contiguous shard of the conditions (sharding.py) and ONE all-reduce of ``(llh, grad[n_problem], n_failed)`` combines
them (NCCL over NVLink on GPUs, gloo in the CPU tests).  Values and gradients are on the model's parameter scale
(``scaled_parameters=True, scaled_gradients=True`` of the reference).  If any condition fails, ``llh`` is NaN and the gradient
is ``None`` — as in the reference, where a failed condition's llh is NaN and ``aggregate_sllh`` returns ``None`` (:301-305).
"""
from __future__ import annotations

import numpy as np
import torch

from .sharding import allreduce_objective, shard_bounds


class ParameterMapping:
    """Array form of PEtab's per-condition ``map_sim_var`` (see module docstring)."""

    def __init__(self, index, value=None):
        self.index = np.ascontiguousarray(index, dtype=np.int64)
        if self.index.ndim != 2:
            raise ValueError("index must be [n_conditions][np]")
        self.value = np.zeros(self.index.shape) if value is None else np.ascontiguousarray(value, dtype=np.float64)
        if self.value.shape != self.index.shape:
            raise ValueError("value must have the shape of index")
        if (self.index < -1).any():
            raise ValueError("index entries must be >= -1")

    def __len__(self):
        return self.index.shape[0]

    def shard(self, lo: int, hi: int) -> "ParameterMapping":
        return ParameterMapping(self.index[lo:hi], self.value[lo:hi])


class FixedParameterMapping(ParameterMapping):
    """Array form of PEtab's ``map_sim_fix`` / ``map_preeq_fix`` (reference petab/v1/_conditions.py: fixed parameters of the
    simulation and of the pre-equilibration condition): ``k[n, i]`` is the UNSCALED value of problem parameter ``index[n, i]``,
    or the number ``value[n, i]`` where ``index[n, i] == -1``.  Fixed parameters carry no gradient."""


def unscale(theta: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """Problem parameters from their parameter scale (0 none, 1 ln, 2 log10; reference include/amici/misc.h:164-195)."""
    return torch.where(scale == 2, torch.pow(10.0, theta), torch.where(scale == 1, torch.exp(theta), theta))


def fill_parameters(theta: torch.Tensor, index: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    """``fill_in_parameters``: p[n, i] = theta[index[n, i]] where mapped, else value[n, i]."""
    return torch.where(index >= 0, theta[index.clamp(min=0)], value)


def aggregate(theta_size: int, index_plist: torch.Tensor, status: torch.Tensor, llh: torch.Tensor, sllh: torch.Tensor | None) -> torch.Tensor:
    """(sum llh, grad[theta_size], n_failed) of one shard; ``index_plist[n, j]`` = problem parameter of sllh[n, j] or -1."""
    out = torch.zeros(theta_size + 2, dtype=torch.float64, device=llh.device)
    ok = status == 0
    out[0] = torch.where(ok, llh, torch.zeros_like(llh)).sum()
    out[-1] = (~ok).sum().to(torch.float64)
    if sllh is not None and sllh.numel():
        mask = (index_plist >= 0) & ok[:, None]
        contrib = torch.where(mask, sllh, torch.zeros_like(sllh))
        out[1:-1].index_add_(0, index_plist.clamp(min=0).reshape(-1), contrib.reshape(-1))
    return out


class ShardedObjective:
    """``llh, grad = objective(theta)`` over all ranks' conditions.

    ``solve`` is the per-shard batched solve: ``solve(p) -> (status, llh, sllh)`` with device tensors (``sllh`` may be ``None``).
    ``from_device_batch`` builds it from a ``DeviceBatch`` (the CUDA engine); tests may inject another one.
    """

    def __init__(self, n_problem: int, mapping: ParameterMapping, plist, solve, device=None, fixed=None, fixed_preeq=None,
                 theta_scale=None):
        self.n_problem = int(n_problem)
        self.device = device or torch.device("cpu")
        # fixed parameters mapped from problem parameters: ``solve`` is then called as solve(p, k, k_preeq) (None = not mapped)
        self.fixed = self.fixed_preeq = None
        for name, fm in (("fixed", fixed), ("fixed_preeq", fixed_preeq)):
            if fm is None:
                continue
            if len(fm) != len(mapping):
                raise ValueError("one fixed-parameter mapping row per condition")
            if fm.index.size and fm.index.max() >= self.n_problem:
                raise ValueError("fixed-parameter mapping refers to a problem parameter >= n_problem")
            setattr(self, name, (torch.as_tensor(fm.index, device=self.device), torch.as_tensor(fm.value, device=self.device)))
        ts = np.zeros(self.n_problem, dtype=np.int64) if theta_scale is None else np.asarray([int(v) for v in theta_scale], dtype=np.int64)
        if ts.shape != (self.n_problem,):
            raise ValueError("theta_scale must have n_problem entries")
        self.theta_scale = torch.as_tensor(ts, device=self.device)
        if mapping.index.size and mapping.index.max() >= self.n_problem:
            raise ValueError("mapping refers to a problem parameter >= n_problem")
        self.index = torch.as_tensor(mapping.index, device=self.device)
        self.value = torch.as_tensor(mapping.value, device=self.device)
        plist = torch.as_tensor(np.asarray(plist, dtype=np.int64), device=self.device)
        self.index_plist = self.index[:, plist] if len(mapping) else self.index[:, :0]
        self.solve = solve
        self.last = None   # (status, llh, sllh) of the last call, per condition of this shard

    @classmethod
    def from_device_batch(cls, db, n_problem: int, mapping: ParameterMapping, fixed=None, fixed_preeq=None, theta_scale=None):
        """Objective over the conditions of a ``DeviceBatch`` (this rank's shard); ``mapping`` has one row per condition.
        ``fixed`` / ``fixed_preeq`` (``FixedParameterMapping``) fill the batch's k / k_preeq from the problem parameters too."""
        if len(mapping) != db.N:
            raise ValueError("one mapping row per condition of the batch")
        if fixed_preeq is not None and "k_preeq" not in db.inputs:
            raise ValueError("the batch has no pre-equilibration condition (k_preeq) to map onto")

        def solve(p, k=None, k_preeq=None):
            db.inputs["p"].copy_(p)
            if k is not None:
                db.inputs["k"].copy_(k)
            if k_preeq is not None:
                db.inputs["k_preeq"].copy_(k_preeq)
            db.solve()
            return db.out["status"], db.out["llh"], db.out.get("sllh")
        plist = db.host_inputs["plist"][: db.q] if db.q else []
        return cls(n_problem, mapping, plist, solve, db.device, fixed=fixed, fixed_preeq=fixed_preeq, theta_scale=theta_scale)

    def __call__(self, theta):
        theta = torch.as_tensor(theta, dtype=torch.float64, device=self.device)
        if theta.shape != (self.n_problem,):
            raise ValueError("theta must have n_problem entries")
        p = fill_parameters(theta, self.index, self.value)
        if self.fixed is None and self.fixed_preeq is None:
            status, llh, sllh = self.solve(p)
        else:
            lin = unscale(theta, self.theta_scale)
            ks = [None if fm is None else fill_parameters(lin, fm[0], fm[1]) for fm in (self.fixed, self.fixed_preeq)]
            status, llh, sllh = self.solve(p, ks[0], ks[1])
        self.last = (status, llh, sllh)
        obj = aggregate(self.n_problem, self.index_plist, status, llh, sllh)
        allreduce_objective(obj)
        o = obj.cpu().numpy()
        if o[-1] > 0:
            return float("nan"), None
        return float(o[0]), (o[1:-1].copy() if sllh is not None else None)


def shard_mapping(mapping: ParameterMapping, world: int, rank: int) -> tuple[int, int, ParameterMapping]:
    lo, hi = shard_bounds(len(mapping), world, rank)
    return lo, hi, mapping.shard(lo, hi)
