"""Device-resident batch: inputs and ReturnData fields live in HBM as torch tensors (torch is used for
device memory, streams and torch.distributed only); the solve is ``ab_solve_device`` on torch's current
stream.  This is the path a sharded objective evaluation uses: each rank keeps its slice of the
conditions resident, and only (sum llh, sum sllh) crosses NVLink (one NCCL all-reduce)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi, load_library, Model, Solver, ExpDataBatch, RDataReporting, _check


def bind_to_gpu_numa(index: int) -> list[int] | None:
    """Pin the calling process to the host cores next to GPU `index` (its PCI device's `local_cpulist`), before any pinned host
    memory is allocated: with first-touch placement the rank's page-locked result buffers then live on the GPU's NUMA node, and
    the eight ranks of one box stream their ReturnData over eight root complexes into their own memory controllers instead of
    contending for one.  Returns the cores, or None if the topology cannot be read (nothing is changed then)."""
    import os
    try:
        pr = torch.cuda.get_device_properties(index)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{dev}/local_cpulist").read().strip()
        cores: set[int] = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cores.update(range(int(lo), int(hi or lo) + 1))
        cores &= os.sched_getaffinity(0)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return sorted(cores)
    except (OSError, ValueError, AttributeError):
        return None


class DeviceBatch:
    def __init__(self, model: Model, solver: Solver, batch: ExpDataBatch, device: torch.device | None = None,
                 fields=None):
        self.lib = load_library()
        self.model, self.solver = model, solver
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.opts = solver._options(model.t0(), model)
        self.opts.posteq = int(len(batch.ts) > 0 and not np.all(np.isfinite(batch.ts)))
        from . import _reinit_mask
        self.opts.reinit_states_sim = _reinit_mask(getattr(batch, "reinit_states_sim", None), model)
        self.opts.reinit_states_presim = _reinit_mask(getattr(batch, "reinit_states_presim", None), model, presim=True)
        d = dict(model.dims(), nz=model.nz, nmaxevent=self.opts.nmaxevent)
        N, nt = len(batch), len(batch.ts)
        plist = np.asarray(model.get_parameter_list(), dtype=np.int32) if batch.plist is None else batch.plist
        pscale = (np.asarray([int(s) for s in model.get_parameter_scale()], dtype=np.int32)
                  if batch.pscale is None else batch.pscale)
        self.N, self.nt, self.nplist = N, nt, len(plist)
        q = self.nplist if self.opts.sensi else 0
        self.q = q
        dev = self.device
        self.host_inputs = dict(ts=batch.ts, plist=plist, pscale=pscale, p=batch.p, k=batch.k)
        if batch.my is not None:
            self.host_inputs["my"] = batch.my
        if batch.sigmay is not None:
            self.host_inputs["sigmay_in"] = batch.sigmay
        for name, arr in (("k_preeq", batch.k_preeq), ("mz", batch.mz), ("sigmaz_in", batch.sigmaz)):
            if arr is not None:
                self.host_inputs[name] = arr
        if getattr(batch, "x0", None) is not None:
            self.host_inputs["x0_in"] = batch.x0
        if getattr(batch, "sx0", None) is not None and self.opts.sensi:
            self.host_inputs["sx0_in"] = batch.sx0
        if getattr(batch, "t_presim", None) is not None and (batch.t_presim > 0).any():
            self.host_inputs["t_presim"] = batch.t_presim
            if batch.k_presim is not None:
                self.host_inputs["k_presim"] = batch.k_presim
        self.inputs = {k_: torch.as_tensor(v).to(dev) for k_, v in self.host_inputs.items()}
        likelihood_only = self.opts.reporting == int(RDataReporting.likelihood)
        lk = {"status", "llh", "chi2", "sllh", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails", "order"}
        self.out = {}
        for name, t, shp in _capi.OUT_FIELDS:
            if fields is not None and name not in fields and name != "status":
                continue
            if likelihood_only and name not in lk:
                continue
            if name in _capi.PREEQ_FIELDS and batch.k_preeq is None:
                continue
            if name in _capi.POSTEQ_FIELDS and np.all(np.isfinite(batch.ts)):
                continue
            shape = shp(d, N, nt, q)
            if int(np.prod(shape)) == 0:
                continue
            self.out[name] = torch.zeros(shape, dtype=torch.float64 if t == "d" else torch.int32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.objective = torch.zeros(q + 2, dtype=torch.float64, device=dev)
        b = _capi.BatchView(N=N, nt=nt, nplist=self.nplist)
        for k_, v in self.inputs.items():
            setattr(b, k_, C.cast(v.data_ptr(), C.POINTER(C.c_double if v.dtype == torch.float64 else C.c_int)))
        for k_, v in self.out.items():
            setattr(b, k_, C.cast(v.data_ptr(), C.POINTER(C.c_double if v.dtype == torch.float64 else C.c_int)))
        self.view = b

    def input_bytes(self) -> int:
        return int(sum(v.numel() * v.element_size() for v in self.inputs.values()))

    def output_bytes(self) -> int:
        return int(sum(v.numel() * v.element_size() for v in self.out.values()))

    def solve(self):
        """Launch the batched solve on torch's current stream (asynchronous)."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        _check(self.lib.ab_solve_device(self.model.name.encode(), C.byref(self.opts), C.byref(self.view),
                                        self.counter.data_ptr(), st))

    def reduce_objective(self) -> torch.Tensor:
        """(sum llh, sum sllh[nplist], n_failed) on the device; all-reduce this tensor across ranks."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        sllh = self.out["sllh"].data_ptr() if "sllh" in self.out else 0
        _check(self.lib.ab_reduce_objective(self.N, self.q, self.out["status"].data_ptr(), self.out["llh"].data_ptr(),
                                            sllh, self.objective.data_ptr(), st))
        return self.objective
