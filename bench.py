#!/usr/bin/env python
"""bench.py — ODE solves/s with forward sensitivities on the JAK-STAT model (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-conditions M]

One "step" = one pass of the batched solve over the synthetic batch (65,536 conditions per GPU, weak
scaling: every rank integrates its own shard; the only collective is one NCCL all-reduce of
(sum llh, sum sllh, n_failed) per step).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ODE solves/sec w/ forward sens (JAK-STAT)"
UNIT = "solves/s"
TS = np.arange(0, 61, 5.0)
RTOL, ATOL = 1e-8, 1e-12
NCU_DRAM_BYTES_PER_CONDITION = (19.332096e6 + 139.997952e6) / 4736   # profiles/r01_v6_ncu_summary.md


def synthetic_inputs(n: int, seed: int, p_nom, k_nom):
    """k x U(0.8,1.25), p x 10^N(0,0.1^2); measurements added later."""
    rng = np.random.default_rng(seed)
    p = np.tile(p_nom, (n, 1)) * 10 ** rng.normal(0, 0.1, (n, len(p_nom)))
    k = np.tile(k_nom, (n, 1)) * rng.uniform(0.8, 1.25, (n, len(k_nom)))
    noise = 0.1 * rng.standard_normal((n, len(TS), 3))
    return p, k, noise


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, row in self.rows:
            if not (t0 <= ts <= t1 + 0.2):
                continue
            f = [c.strip() for c in row.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops_per_solve(nst, nni, netf, ncfn, nx=9, s=17, nt=13, ny=3):
    """FLOPs of one solve counted from the REFERENCE's solver counters."""
    zJ, Ff, FJ, FP = 18, 45, 12, 30   # JAK-STAT: nnz(J), flops of xdot(+spline), jac, dxdotdp column set
    n = nx
    F_it = Ff + (FJ + FP + s * (2 * zJ + n)) + 4 * n * (1 + s) + 2 * n * n * (1 + s) + 3 * n * (1 + s)
    F_st = 39 * n * (1 + s)
    nsetups = nst / 6.0   # the reference driver does not export nsetups; ~1 per 6 steps measured in DESIGN.md
    F_ls = FJ + n * n + (2.0 / 3.0) * n ** 3
    F_out = 2 * ny * n * s + 2 * n * s + 3 * ny * s * s
    return nni * F_it + (nst + netf + ncfn) * F_st + nsetups * F_ls + nt * F_out


def cpu_reference_run(p, k, my, sig, threads=None):
    from oracle.refsim import RefModel
    ref = RefModel("jakstat_adjoint")
    threads = threads or ref.max_threads()
    r = ref.run(TS, p, k, my=my, sigmay=sig, rtol=RTOL, atol=ATOL, sensi_order=1, num_threads=threads,
                fields=["status", "llh", "sllh", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails"])
    return r, threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.refsim import RefModel, available
    if not available("jakstat_adjoint"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_jakstat_adjoint.so not built"}))
        return
    ref = RefModel("jakstat_adjoint")
    cores = ref.max_threads()
    sample = int(min(args.n_conditions, max(256, 16 * cores)))
    p, k, noise = synthetic_inputs(sample, 2, ref.p_nominal, ref.k_nominal)
    y = ref.run(TS, p, k, sensi_order=0, num_threads=cores, fields=["y"])["y"]
    my, sig = y + noise, np.full_like(y, 0.1)
    times = []
    for i in range(args.warmup + args.steps):
        r, _ = cpu_reference_run(p, k, my, sig, cores)
        if i >= args.warmup:
            times.append(r["wall_seconds"])
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"model_jakstat_adjoint_py FSA 17 params, {sample} conditions/step (bounded sample of 65536), "
                                   f"13 timepoints, rtol {RTOL} atol {ATOL}, amici::run_simulations OpenMP {cores} threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{sample} conditions per step, unmodified reference compiled from /root/reference (oracle/_ref)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-conditions", type=int, default=65536, help="conditions per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import amici_b200 as ab
    from amici_b200.device import DeviceBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ab.load_library()

    model = ab.get_model("jakstat_adjoint")
    model.set_timepoints(TS)
    solver = model.create_solver()
    solver.set_relative_tolerance(RTOL)
    solver.set_absolute_tolerance(ATOL)
    solver.set_sensitivity_order(ab.SensitivityOrder.first)
    if os.environ.get("BENCH_NPLIST"):      # experiment only: sensitivities w.r.t. the first parameters only
        model.set_parameter_list(list(range(int(os.environ["BENCH_NPLIST"]))))
    if os.environ.get("BENCH_NO_STLDET"):   # experiment only: upper bound of what cvBDFStab costs (not the reference's default)
        solver.set_stability_limit_flag(False)
    n = args.n_conditions
    p, k, noise = synthetic_inputs(n, 2 + 1000 * rank, np.array(model.get_free_parameters()), np.array(model.get_fixed_parameters()))
    # noise-free observables of every condition (one un-timed solve without sensitivities), then add noise
    s0 = model.create_solver()
    s0.set_relative_tolerance(RTOL)
    s0.set_absolute_tolerance(ATOL)
    pre = DeviceBatch(model, s0, ab.ExpDataBatch(TS, p, k), dev, fields=["y"])
    pre.solve()
    torch.cuda.synchronize()
    my = pre.out["y"].cpu().numpy() + noise
    del pre
    sig = np.full_like(my, 0.1)
    batch = ab.ExpDataBatch(TS, p, k, my=my, sigmay=sig)

    # ---------------- device-resident throughput (value) ----------------
    db = DeviceBatch(model, solver, batch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    obj = None

    def step():
        nonlocal obj
        db.solve()
        obj = db.reduce_objective()
        if world > 1:
            dist.all_reduce(obj)

    for _ in range(args.warmup):
        step()
        flush.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.ab_kernel_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    torch.cuda.synchronize()
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        flush.zero_()           # L2 flush between timed iterations (outside the event pair)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = lib.ab_kernel_launch_count() - launches0
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    status = db.out["status"].cpu().numpy()
    n_failed = int((status != 0).sum())
    objective = obj.cpu().numpy()

    # ---------------- dominant kernel alone (roofline) ----------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        kev[i][0].record()
        db.solve()
        kev[i][1].record()
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    in_b, out_b = db.input_bytes(), db.output_bytes()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"model_jakstat_adjoint_py FSA 17 params, {n} synthetic ExpData conditions per GPU per step, "
                                   f"13 timepoints 0..60, rtol {RTOL} atol {ATOL}, RDataReporting.full",
                       "parallelism": f"shard{world}" if world > 1 else "single",
                       "l2": "explicit 256 MiB flush between timed steps; per-step traffic (2.4 GB of outputs) exceeds the 126 MB L2",
                       "n_failed": n_failed, "sum_llh": float(objective[0])},
            "clocks": clocks, "gpu_launches": int(launches)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        ach = (in_b + out_b) / (kernel_ms * 1e-3) / 1e9
        # DRAM bytes of this kernel from the ncu --set full capture under profiles/ (r01_v6_ncu_summary.md; 4736-condition launch), scaled to this launch's condition count
        traffic = NCU_DRAM_BYTES_PER_CONDITION * n
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                            "traffic": traffic, "traffic_source": "profiles/r01_v6_ncu_summary.md (dram__bytes_read+write per condition of a 4736-condition capture) x conditions per launch",
                            "kernel": "bdf_solve_kernel<jakstat_adjoint,32,1,256,1>", "kernel_ms": kernel_ms,
                            "algorithmic_bytes_per_launch": in_b + out_b, "peak_source": peak_src,
                            "note": "the solve is FP64-latency bound, not HBM bound (state lives in shared memory/registers); see roofline_fp64"}
        import ctypes as C
        tf = C.c_double(0.0)
        lib.ab_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
        lib.ab_measure_fp64_peak(C.byref(tf))
        line["roofline_fp64"] = {"peak": tf.value, "unit": "TFLOP/s", "peak_source": "measured in this run (ab_measure_fp64_peak: dependent DFMA chains on all SMs)"}

    # ---------------- end to end through the public API with host buffers ----------------
    if not args.no_e2e:
        del db
        torch.cuda.empty_cache()
        e2e_t = []
        for i in range(max(1, min(args.warmup, 2)) + args.steps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = ab.run_simulations(model, solver, batch)       # host numpy in, host (pinned) numpy out
            llh_sum = float(res["llh"][res["status"] == 0].sum())  # reads the step's result on the host
            t1 = time.perf_counter()
            if i >= max(1, min(args.warmup, 2)):
                e2e_t.append(t1 - t0)
            h2d = int(batch.ts.nbytes + batch.p.nbytes + batch.k.nbytes + batch.my.nbytes + batch.sigmay.nbytes + 4 * (17 + 17))
            d2h = int(sum(v.nbytes for v in res.fields.values()))
            res.release()   # hand the pinned output buffers back for the next step
            del res
        if os.environ.get("BENCH_DEBUG"):
            print("e2e times", e2e_t, file=sys.stderr)
        e2e_s = float(np.mean(e2e_t))
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        line["e2e"] = {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_s * 1e3, "api": "amici_b200.run_simulations(model, solver, ExpDataBatch) -> ab_run_batch (host buffers)"}

    # ---------------- CPU baseline: the unmodified reference on the host cores (rank 0, N=1 only) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.refsim import available
        if available("jakstat_adjoint"):
            cores = os.cpu_count() or 1
            sample = int(min(n, max(256, 16 * cores)))
            r, thr = cpu_reference_run(p[:sample], k[:sample], my[:sample], sig[:sample])
            r, thr = cpu_reference_run(p[:sample], k[:sample], my[:sample], sig[:sample])
            cpu_val = sample / r["wall_seconds"]
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": thr, "kind": "reference",
                                    "sample": f"first {sample} conditions of the same batch, amici::run_simulations with {thr} OpenMP threads (second of two runs)"}
            fl = algorithmic_flops_per_solve(float(r["numsteps"][:, -1].mean()), float(r["numrhsevals"][:, -1].mean()),
                                             float(r["numerrtestfails"][:, -1].mean()), float(r["numnonlinsolvconvfails"][:, -1].mean()))
            tfs = fl * n / (kernel_ms * 1e-3) / 1e12
            line["roofline_fp64"].update({"achieved": tfs, "frac": tfs / line["roofline_fp64"]["peak"] if line["roofline_fp64"]["peak"] else None,
                                          "algorithmic_flops_per_solve": fl, "counters": "mean of the reference's nst/nrhs/netf/ncfn on the CPU sample"})
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
