#!/usr/bin/env python
"""bench.py — ODE solves/s of the batched BDF engine on the BASELINE.json configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--n-conditions M]

--config (default jakstat = BASELINE.json configs[1], the configuration the metric is quoted on):
  jakstat      model_jakstat_adjoint_py, 17 forward sensitivities, 65,536 synthetic ExpData conditions
  robertson    Robertson (ODE form), forward simulation, 4096 log-uniform parameter draws            (configs[0])
  robertson_dae  the shipped model_robertson_py (index-1 DAE): engine DAE path vs the reference's IDAS, same draws          (configs[0])
  calvetti     the shipped model_calvetti_py (index-1 DAE with switches), 16,384 draws of its constants: DAE path vs IDAS
  steadystate  model_steadystate_py, pre-equilibration (Newton, simulation if it fails) + FSA, 65,536  (configs[2])
  events       model_events_py, Heaviside events + root finding + FSA, 65,536 parameter samples        (configs[3])
  neuron       model_neuron_py, spike events (divergent step control) + FSA, 65,536 samples            (configs[3])
  massaction   synthetic mass-action network 100 species / 300 reactions, FSA w.r.t. all 300 rate constants, all-reduced
               llh / sllh; 2368 samples per GPU per step (a bounded sample of the 1M-sample config)    (configs[4])

One "step" = one pass of the batched solve over the synthetic batch (weak scaling: every rank integrates its own shard; the
only collective is one NCCL all-reduce of (sum llh, sum sllh, n_failed) per step).  Prints ONE JSON line on rank 0.
`--impl reference` times the UNMODIFIED reference (oracle/_ref, compiled from /root/reference) on the host cores.
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

UNIT = "solves/s"
RTOL, ATOL = 1e-8, 1e-12
FP64_DATASHEET_TFLOPS = 40.0   # NVIDIA B200 datasheet FP64 (vector) figure, printed next to the measured peak


# --------------------------------------------------------------------------------------------------------------------
# workloads: seeded synthetic inputs, shared by the engine and the reference arm
# --------------------------------------------------------------------------------------------------------------------
def _wl_jakstat(n, seed, p0, k0):
    rng = np.random.default_rng(seed)
    p = np.tile(p0, (n, 1)) * 10 ** rng.normal(0, 0.1, (n, len(p0)))
    k = np.tile(k0, (n, 1)) * rng.uniform(0.8, 1.25, (n, len(k0)))
    return dict(p=p, k=k, noise=0.1 * rng.standard_normal((n, 13, 3)))


def _wl_robertson(n, seed, p0, k0):
    rng = np.random.default_rng(seed)
    return dict(p=np.tile(p0, (n, 1)) * 10 ** rng.uniform(-1, 1, (n, 3)), k=np.full((n, 1), 0.9), noise=None)


def _wl_calvetti(n, seed, p0, k0):
    rng = np.random.default_rng(seed)
    return dict(p=np.zeros((n, 0)), k=np.tile(k0, (n, 1)) * rng.uniform(0.9, 1.1, (n, len(k0))), noise=None)


def _wl_steadystate(n, seed, p0, k0):
    rng = np.random.default_rng(seed)
    p = np.tile(p0, (n, 1)) * 10 ** rng.normal(0, 0.1, (n, 5))
    k = np.tile(k0, (n, 1)) * rng.uniform(0.5, 2, (n, 4))
    kp = np.tile(k0, (n, 1)) * rng.uniform(0.5, 2, (n, 4))
    return dict(p=p, k=k, k_preeq=kp, noise=0.1 * rng.standard_normal((n, 11, 3)))


def _wl_events(n, seed, p0, k0):
    rng = np.random.default_rng(seed)
    return dict(p=np.tile(p0, (n, 1)) * 10 ** rng.normal(0, 0.1, (n, len(p0))), k=np.tile(k0, (n, 1)),
                noise=0.1 * rng.standard_normal((n, 21, 1)))


def _wl_massaction(n, seed, p0, k0):
    rng = np.random.default_rng(seed)
    return dict(p=np.tile(p0, (n, 1)) * 10 ** rng.uniform(-1, 1, (n, len(p0))), k=np.ones((n, 1)), noise="uniform")


# n / s / zJ / Ff / FJ / FP / ny: state dimension, sensitivities, nnz(J), flops of xdot, of the Jacobian, of dxdot/dp
CONFIGS = {
    "jakstat": dict(model="jakstat_adjoint", ref="jakstat_adjoint", n=65536, seed=2, ts=np.arange(0, 61, 5.0), sensi=True, inputs=_wl_jakstat,
                    label="model_jakstat_adjoint_py FSA 17 params", metric="ODE solves/sec w/ forward sens (JAK-STAT)",
                    flops=dict(n=9, s=17, zJ=18, Ff=45, FJ=12, FP=30, ny=3), baseline_config=1),
    # the same workload with SensitivityMethod::adjoint on both arms (amici_b200/csrc/adjoint.cuh; reference src/backwardproblem.cpp):
    # llh + gradient only (the adjoint method has no sx / sy), likelihood reporting
    "jakstat_asa": dict(model="jakstat_adjoint", ref="jakstat_adjoint", n=65536, seed=2, ts=np.arange(0, 61, 5.0), sensi=True, inputs=_wl_jakstat,
                        label="model_jakstat_adjoint_py ADJOINT sensitivities 17 params, likelihood reporting",
                        metric="ODE solves/sec w/ adjoint sens (JAK-STAT)", adjoint=True, likelihood_only=True,
                        flops=dict(n=9, s=17, zJ=18, Ff=45, FJ=12, FP=30, ny=3), baseline_config=1),
    "robertson": dict(model="robertson_ode", ref="robertson_ode", n=4096, seed=1, ts=np.concatenate([[0.0], 4 * np.logspace(-6, 6, 50)]),
                      sensi=False, inputs=_wl_robertson, label="Robertson (ODE form, CVODES-equivalent BDF) forward simulation, log-uniform draws",
                      metric="ODE solves/sec, forward simulation (Robertson)", flops=dict(n=3, s=0, zJ=7, Ff=12, FJ=8, FP=0, ny=3),
                      baseline_config=0),
    # BASELINE configs[0] as literally named: the SHIPPED model_robertson_py (index-1 DAE) — engine: the DAE path (csrc/dae.cuh, one
    # thread per condition), reference arm: the shipped model through IDAS
    "robertson_dae": dict(model="robertson", ref="robertson", n=4096, seed=1, ts=np.concatenate([[0.0], 4 * np.logspace(-6, 6, 50)]),
                          sensi=False, inputs=_wl_robertson, label="model_robertson_py (shipped index-1 DAE; engine: mass-matrix BDF, reference: IDAS) forward simulation, log-uniform draws",
                          metric="DAE solves/sec, forward simulation (Robertson DAE)", flops=dict(n=3, s=0, zJ=7, Ff=12, FJ=8, FP=0, ny=3),
                          baseline_config=0),
    # the other shipped DAE model: three differential + three algebraic states, switches at t = 10 and t = 12 (not a BASELINE config)
    "calvetti": dict(model="calvetti", ref="calvetti", n=16384, seed=6, ts=np.linspace(0, 20, 41), sensi=False, inputs=_wl_calvetti,
                     label="model_calvetti_py (shipped index-1 DAE with time-triggered switches; reference: IDAS) forward simulation, constants x U(0.9, 1.1)",
                     metric="DAE solves/sec, forward simulation (Calvetti)", flops=dict(n=6, s=0, zJ=23, Ff=80, FJ=120, FP=0, ny=6),
                     baseline_config=0),
    "steadystate": dict(model="steadystate", ref="steadystate", n=65536, seed=3, ts=np.linspace(0, 10, 11), sensi=True, inputs=_wl_steadystate,
                        preeq=True, label="model_steadystate_py pre-equilibration (Newton, simulation if it fails) + FSA 5 params",
                        metric="ODE solves/sec w/ forward sens + pre-equilibration (steadystate)",
                        flops=dict(n=3, s=5, zJ=9, Ff=20, FJ=12, FP=11, ny=3), baseline_config=2),
    "events": dict(model="events", ref="events", n=65536, seed=4, ts=np.linspace(0, 10, 21), sensi=True, inputs=_wl_events, nmaxevent=2,
                   label="model_events_py Heaviside events + root finding + FSA 4 params", metric="ODE solves/sec w/ forward sens + events (events)",
                   flops=dict(n=3, s=4, zJ=5, Ff=14, FJ=6, FP=8, ny=1), baseline_config=3),
    "neuron": dict(model="neuron", ref="neuron", n=65536, seed=4, ts=np.linspace(0, 100, 21), sensi=True, inputs=_wl_events, nmaxevent=10,
                   label="model_neuron_py spike events (divergent step control) + FSA 4 params", metric="ODE solves/sec w/ forward sens + events (neuron)",
                   flops=dict(n=2, s=4, zJ=4, Ff=10, FJ=4, FP=4, ny=1), baseline_config=3),
    "massaction": dict(model="massaction", ref="massaction", n=2368, seed=5, ts=np.linspace(0, 10, 11), sensi=True, inputs=_wl_massaction,
                       label="synthetic mass-action network 100 species / 300 reactions, FSA 300 params, likelihood reporting",
                       metric="ODE solves/sec w/ forward sens (mass-action 100x300)", flops=dict(n=100, s=300, zJ=500, Ff=900, FJ=1000, FP=600, ny=10),
                       baseline_config=4, likelihood_only=True, hbm_bytes_per_step=5.2e6, cpu_sample=48),
}


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


def algorithmic_flops_per_solve(c, nt, nst, nni, netf, ncfn):
    """FLOPs of one solve counted from solver counters (the reference's where available).  The number of
    linear-solver setups is not exported by the reference's ReturnData; the engine's own counter on the golden batches is
    one setup per 5.9 steps (tests/cpu_emul), F_ls is < 1 % of the total either way."""
    f = c["flops"]
    n, s = f["n"], f["s"]
    F_it = f["Ff"] + (f["FJ"] + f["FP"] + s * (2 * f["zJ"] + n) if s else 0) + 4 * n * (1 + s) + 2 * n * n * (1 + s) + 3 * n * (1 + s)
    F_st = 39 * n * (1 + s)
    F_ls = f["FJ"] + n * n + (2.0 / 3.0) * n ** 3
    F_out = 2 * f["ny"] * n * s + 2 * n * s + 3 * f["ny"] * s * s
    return nni * F_it + (nst + netf + ncfn) * F_st + (nst / 5.9) * F_ls + nt * F_out


def reference_kwargs(c):
    kw = dict(rtol=RTOL, atol=ATOL, sensi_order=1 if c["sensi"] else 0)
    if c.get("preeq"):
        kw.update(newton_maxsteps=50, ss_mode=2, ss_sensi_mode=0)
    if "nmaxevent" in c:
        kw["nmaxevent"] = c["nmaxevent"]
    if c.get("likelihood_only"):
        kw["reporting"] = 1
    if c.get("adjoint"):
        kw["sensi_meth"] = 2
        kw["reporting"] = 2   # RDataReporting::likelihood
    return kw


def reference_measurements(ref, c, w):
    """noise-free observables of the reference + noise (the engine arm does the same with its own observables)"""
    if w["noise"] is None:
        return None, None
    if isinstance(w["noise"], str):
        rng = np.random.default_rng(12345)
        my = rng.uniform(0.5, 1.5, (len(w["p"]), len(c["ts"]), ref.dims.ny))
        return my, np.full_like(my, 0.1)
    kw = {k_: v for k_, v in reference_kwargs(c).items() if k_ not in ("sensi_order", "reporting")}
    y = ref.run(c["ts"], w["p"], w["k"], k_preeq=w.get("k_preeq"), sensi_order=0, num_threads=ref.max_threads(), fields=["y"], **kw)["y"]
    my = y + w["noise"]
    return my, np.full_like(my, 0.1)


COUNTERS = ["status", "llh", "numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails"]


def time_reference(ref, c, w, my, sig, threads, reps):
    times, r = [], None
    for _ in range(reps):
        r = ref.run(c["ts"], w["p"], w["k"], k_preeq=w.get("k_preeq"), my=my, sigmay=sig, num_threads=threads, fields=COUNTERS,
                    **reference_kwargs(c))
        times.append(r["wall_seconds"])
    return times, r


def run_reference_arm(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.refsim import RefModel, available
    if not available(c["ref"]):
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libref_{c['ref']}.so not built"}))
        return
    ref = RefModel(c["ref"])
    cores = ref.max_threads()
    # a fixed subset of >= 4096 conditions of the batch (the mass-action config: ~13 s per condition-core)
    sample = int(min(args.n_conditions or c["n"], c.get("cpu_sample", 4096)))
    w = c["inputs"](sample, c["seed"], ref.p_nominal, ref.k_nominal)
    my, sig = reference_measurements(ref, c, w)
    times, r = time_reference(ref, c, w, my, sig, cores, args.warmup + args.steps)
    timed = times[args.warmup:]
    best, mean = float(np.min(timed)), float(np.mean(timed))
    # one-thread figure on a smaller subset (BASELINE.md §3.3 asks for T = 1 next to T = all)
    s1 = max(8, min(sample, sample // (2 * cores)))
    w1 = {k_: (v[:s1] if isinstance(v, np.ndarray) else v) for k_, v in w.items()}
    t1, _ = time_reference(ref, c, w1, None if my is None else my[:s1], None if sig is None else sig[:s1], 1, 1)
    val = sample / best
    line = {"impl": "reference", "metric": c["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * best, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{c['label']}, {sample} conditions/step (fixed subset of the {c['n']}-condition batch, BASELINE.md §3.3), "
                                   f"{len(c['ts'])} timepoints, rtol {RTOL} atol {ATOL}, amici::run_simulations OpenMP {cores} threads",
                       "baseline_config": c["baseline_config"], "timing": f"value = best of the {len(timed)} timed steps; mean {sample / mean:.1f} solves/s",
                       "n_failed": int((r["status"] != 0).sum())},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "one_thread": {"value": s1 / t1[0], "sample": s1},
                             "sample": f"{sample} conditions per step, unmodified reference compiled from /root/reference (oracle/_ref), best of {len(timed)}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def build_engine_objects(ab, c, n, rank):
    model = ab.get_model(c["model"])
    model.set_timepoints(c["ts"])
    if "nmaxevent" in c:
        model.set_n_max_event(c["nmaxevent"])
    if c.get("preeq"):
        model.set_steady_state_computation_mode(ab.SteadyStateComputationMode.integrateIfNewtonFails)
        model.set_steady_state_sensitivity_mode(ab.SteadyStateSensitivityMode.newtonOnly)
    if os.environ.get("BENCH_NPLIST"):      # experiment only: sensitivities w.r.t. the first parameters only
        model.set_parameter_list(list(range(int(os.environ["BENCH_NPLIST"]))))

    def make_solver(sensi):
        s = model.create_solver()
        s.set_relative_tolerance(RTOL)
        s.set_absolute_tolerance(ATOL)
        if sensi:
            s.set_sensitivity_order(ab.SensitivityOrder.first)
        if c.get("preeq"):
            s.set_newton_max_steps(50)
        if c.get("likelihood_only"):
            s.set_return_data_reporting_mode(ab.RDataReporting.likelihood)
        if c.get("adjoint") and sensi:
            s.set_sensitivity_method(ab.SensitivityMethod.adjoint)
        if os.environ.get("BENCH_NO_STLDET"):   # experiment only: what cvBDFStab costs (not the reference's default)
            s.set_stability_limit_flag(False)
        return s
    w = c["inputs"](n, c["seed"] + 1000 * rank, np.array(model.get_free_parameters()), np.array(model.get_fixed_parameters()))
    return model, make_solver, w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="jakstat", choices=sorted(CONFIGS))
    ap.add_argument("--n-conditions", type=int, default=0, help="conditions per GPU per step (default: the config's BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    c = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference_arm(args, c)

    import torch
    import torch.distributed as dist
    import amici_b200 as ab
    from amici_b200.device import DeviceBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = None
    if world > 1 and not os.environ.get("BENCH_NO_NUMA_BIND"):
        from amici_b200.device import bind_to_gpu_numa
        numa_cores = bind_to_gpu_numa(local)   # before the pinned result buffers of the end-to-end leg are allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ab.load_library()

    n = args.n_conditions or c["n"]
    model, make_solver, w = build_engine_objects(ab, c, n, rank)
    solver = make_solver(c["sensi"])
    p, k, kp = w["p"], w["k"], w.get("k_preeq")
    my = sig = None
    if isinstance(w["noise"], str):
        rng = np.random.default_rng(12345 + rank)
        my = rng.uniform(0.5, 1.5, (n, len(c["ts"]), model.ny))
    elif w["noise"] is not None:
        # noise-free observables of every condition (one un-timed solve without sensitivities), then add noise
        pre_solver = make_solver(False)
        pre_solver.set_return_data_reporting_mode(ab.RDataReporting.full)   # the observables are needed here, whatever the timed mode
        pre = DeviceBatch(model, pre_solver, ab.ExpDataBatch(c["ts"], p, k, k_preeq=kp), dev, fields=["y"])
        pre.solve()
        torch.cuda.synchronize()
        my = pre.out["y"].cpu().numpy() + w["noise"]
        del pre
    if my is not None:
        sig = np.full_like(my, 0.1)
    batch = ab.ExpDataBatch(c["ts"], p, k, my=my, sigmay=sig, k_preeq=kp)

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
    own = {f: float(db.out[f][:, -1].double().mean()) for f in ("numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails") if f in db.out}

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

    line = {"metric": c["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{c['label']}, {n} synthetic conditions per GPU per step, {len(c['ts'])} timepoints, rtol {RTOL} atol {ATOL}, "
                                   + ("RDataReporting.likelihood" if c.get("likelihood_only") else "RDataReporting.full"),
                       "baseline_config": c["baseline_config"], "name": args.config,
                       "parallelism": f"shard{world}" if world > 1 else "single",
                       "numa_bind": (f"rank 0 on {len(numa_cores)} cores next to its GPU" if numa_cores else "off"),
                       "l2": "explicit 256 MiB flush between timed steps",
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
        # DRAM bytes of the kernel: per-condition figure of the committed ncu --set full capture of this kernel (written by
        # tools/ncu_summary.py traffic), scaled to this launch; null when no capture of this config is committed
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.config)
            if tr:
                traffic, traffic_src = tr["dram_bytes_per_condition"] * n, tr["source"]
        except OSError:
            pass
        import ctypes as C
        tf = C.c_double(0.0)
        lib.ab_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
        lib.ab_measure_fp64_peak(C.byref(tf))
        hbm_ach = (in_b + out_b) / (kernel_ms * 1e-3) / 1e9
        kname = lib.ab_last_kernel_name().decode() if hasattr(lib, "ab_last_kernel_name") else c["model"]
        hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "traffic": traffic,
               "traffic_source": traffic_src, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": in_b + out_b, "peak_source": peak_src}
        line["_counters"] = own
        line["_fp64_peak"] = tf.value
        line["_hbm"] = hbm
        line["_kernel"] = kname

    # ---------------- end to end through the public API with host buffers ----------------
    if not args.no_e2e:
        del db
        torch.cuda.empty_cache()
        e2e_t = []
        nwarm = max(1, min(args.warmup, 2))
        for i in range(nwarm + args.steps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = ab.run_simulations(model, solver, batch)       # host numpy in, host (pinned) numpy out
            llh_sum = float(res["llh"][res["status"] == 0].sum())  # reads the step's result on the host
            t1 = time.perf_counter()
            if i >= nwarm:
                e2e_t.append(t1 - t0)
            h2d = int(sum(a.nbytes for a in (batch.ts, batch.p, batch.k, batch.my, batch.sigmay, batch.k_preeq) if a is not None)
                      + 4 * 2 * model.np())
            d2h = int(sum(v.nbytes for v in res.fields.values()))
            res.release()   # hand the pinned output buffers back for the next step
            del res
        e2e_s = float(np.mean(e2e_t))
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        line["e2e"] = {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_s * 1e3, "api": "amici_b200.run_simulations(model, solver, ExpDataBatch) -> ab_run_batch (host buffers)"}

    # ---------------- CPU baseline: the unmodified reference on the host cores (rank 0, N=1 only) ----------------
    counters, counters_src = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.refsim import RefModel, available
        if available(c["ref"]):
            ref = RefModel(c["ref"])
            cores = ref.max_threads()
            sample = int(min(n, c.get("cpu_sample", 4096)))
            ws = {k_: (v[:sample] if isinstance(v, np.ndarray) else v) for k_, v in w.items()}
            mys, sigs = (None, None) if my is None else (my[:sample], sig[:sample])
            times, r = time_reference(ref, c, ws, mys, sigs, cores, 3)
            s1 = max(8, min(sample, sample // (2 * cores)))
            w1 = {k_: (v[:s1] if isinstance(v, np.ndarray) else v) for k_, v in ws.items()}
            t1, _ = time_reference(ref, c, w1, None if mys is None else mys[:s1], None if sigs is None else sigs[:s1], 1, 1)
            line["cpu_baseline"] = {"value": sample / min(times), "unit": UNIT, "cores": cores, "kind": "reference",
                                    "one_thread": {"value": s1 / t1[0], "sample": s1},
                                    "sample": f"first {sample} conditions of the same batch, amici::run_simulations with {cores} OpenMP threads, best of 3 "
                                              f"(mean {sample / float(np.mean(times)):.1f}); one thread on the first {s1}"}
            counters = {f: float(r[f][:, -1].mean()) for f in ("numsteps", "numrhsevals", "numerrtestfails", "numnonlinsolvconvfails")}
            counters_src = f"mean of the reference's nst/nrhs/netf/ncfn on the CPU sample ({sample} conditions)"
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    if rank == 0:
        own = line.pop("_counters")
        peak = line.pop("_fp64_peak")
        hbm = line.pop("_hbm")
        kname = line.pop("_kernel")
        if counters is None:
            counters, counters_src = own, "the engine's own counters (the reference was not run)"
        fl = algorithmic_flops_per_solve(c, len(c["ts"]), counters["numsteps"], counters["numrhsevals"], counters["numerrtestfails"],
                                         counters["numnonlinsolvconvfails"])
        tfs = fl * n / (kernel_ms * 1e-3) / 1e12
        fp64 = {"bound": "fp64", "achieved": tfs, "peak": peak, "unit": "TFLOP/s", "frac": tfs / peak if peak else None,
                "traffic": hbm["traffic"], "traffic_source": hbm["traffic_source"], "kernel": kname, "kernel_ms": kernel_ms,
                "algorithmic_flops_per_solve": fl, "counters": counters_src,
                "peak_source": "SELF-MEASURED in this run (ab_measure_fp64_peak: independent dependent-DFMA chains on all SMs); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "datasheet_peak": FP64_DATASHEET_TFLOPS, "frac_of_datasheet": tfs / FP64_DATASHEET_TFLOPS}
        if "hbm_bytes_per_step" in c:
            # block solver: the per-condition history is HBM-resident and streamed every step
            ach = c["hbm_bytes_per_step"] * counters["numsteps"] * n / (kernel_ms * 1e-3) / 1e9
            hbm.update(achieved=ach, frac=ach / hbm["peak"], algorithmic_bytes_per_launch=c["hbm_bytes_per_step"] * counters["numsteps"] * n,
                       note="algorithmic bytes = 5.2 MB per BDF step and condition (DESIGN.md §2.8) x steps")
            line["roofline"], line["roofline_fp64"] = hbm, fp64
        else:
            hbm["note"] = "inputs + ReturnData only: the solve keeps its state in shared memory / registers and is not HBM bound"
            line["roofline"], line["roofline_hbm"] = fp64, hbm
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
