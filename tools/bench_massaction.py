#!/usr/bin/env python
"""BASELINE.json configs[4]: synthetic mass-action network (100 species / 300 reactions), forward sensitivities, llh/sllh
reduced on the device and all-reduced over NCCL.  Same conventions as bench.py (one JSON line on rank 0); this is a
secondary workload, the headline bench stays bench.py (JAK-STAT).

  python tools/bench_massaction.py [--gpus N] [--steps K] [--warmup W] [--n-conditions M] [--nplist Q]
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_massaction.py --gpus N
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--n-conditions", type=int, default=2368, help="conditions per GPU per step")
    ap.add_argument("--nplist", type=int, default=300)
    ap.add_argument("--model", default="massaction")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import amici_b200 as ab
    from amici_b200.codegen.massaction import network
    from amici_b200.device import DeviceBatch
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = network(args.model)
    npl = min(args.nplist, net["nr"])
    model = ab.get_model(args.model)
    ts = np.linspace(0, 10, 11)
    model.set_timepoints(ts)
    model.set_parameter_list(list(range(npl)))
    solver = model.create_solver()
    solver.set_relative_tolerance(1e-8)
    solver.set_absolute_tolerance(1e-12)
    solver.set_sensitivity_order(ab.SensitivityOrder.first)
    solver.set_return_data_reporting_mode(ab.RDataReporting.likelihood)
    n = args.n_conditions
    rng = np.random.default_rng(5 + 1000 * rank)
    p = np.tile(net["p_nominal"], (n, 1)) * 10 ** rng.uniform(-1, 1, (n, net["nr"]))
    k = np.ones((n, 1))
    my = rng.uniform(0.5, 1.5, (n, len(ts), net["ny"]))
    sig = np.full_like(my, 0.1)
    batch = ab.ExpDataBatch(ts, p, k, my=my, sigmay=sig)
    db = DeviceBatch(model, solver, batch, dev)
    obj = None

    def step():
        nonlocal obj
        db.solve()
        obj = db.reduce_objective()
        if world > 1:
            dist.all_reduce(obj)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
    torch.cuda.synchronize()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms = total_ms / args.steps
    status = db.out["status"].cpu().numpy()
    line = {"metric": f"ODE solves/sec w/ forward sens ({args.model}: {net['nx']} species / {net['nr']} reactions, {npl} sensitivities)",
            "value": world * n / (ms * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.model} FSA {npl} params, {n} conditions per GPU per step, rate constants nominal x 10^U(-1,1), "
                                   "11 timepoints 0..10, rtol 1e-8 atol 1e-12, RDataReporting.likelihood (llh, sllh reduced on device"
                                   + (", NCCL all-reduce" if world > 1 else "") + ")",
                       "n_failed": int((status != 0).sum()), "sum_llh": float(obj.cpu().numpy()[0])}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.refsim import RefModel, available
        if available(args.model):
            ref = RefModel(args.model)
            cores = ref.max_threads()
            sample = min(n, cores)
            t0 = time.perf_counter()
            r = ref.run(ts, p[:sample], k[:sample], my=my[:sample], sigmay=sig[:sample], plist=list(range(npl)), sensi_order=1,
                        rtol=1e-8, atol=1e-12, num_threads=cores, fields=["status", "llh", "sllh"])
            wall = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / wall, "unit": "solves/s", "cores": cores, "kind": "reference",
                                    "sample": f"first {sample} conditions of the same batch, amici::run_simulations with {cores} OpenMP threads"}
            llh_gpu = db.out["llh"].cpu().numpy()[:sample]
            line["config"]["max_rel_llh_dev_vs_reference_on_sample"] = float(np.max(np.abs(llh_gpu - r["llh"]) / np.abs(r["llh"])))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
