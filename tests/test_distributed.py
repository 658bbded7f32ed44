"""world_size-2 gloo test of the sharded objective: shard -> local solve -> local reduction -> one all-reduce."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from amici_b200 import _capi
    from amici_b200.sharding import allreduce_objective, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "steadystate.npz"))
    lo, hi = shard_bounds(len(g["p"]), world, rank)
    lib = C.CDLL(os.path.join(ROOT, "tests", "cpu_emul", "libemul.so"))
    dims = dict(nx=3, np=5, nk=4, ny=3)
    b, out, keep = _capi.make_batch(dims, g["ts"], g["p"][lo:hi], g["k"][lo:hi], my=g["my"][lo:hi], sigmay=g["sigmay_in"][lo:hi])
    o = _capi.default_options(rtol=1e-8, atol=1e-12, sensi=1)
    assert lib.emul_run(b"steadystate", C.byref(o), C.byref(b)) == 0
    ok = out["status"] == 0
    obj = torch.tensor(np.concatenate([[out["llh"][ok].sum()], out["sllh"][ok].sum(0), [float((~ok).sum())]]))
    allreduce_objective(obj)
    q.put((rank, lo, hi, obj.numpy()))
    dist.destroy_process_group()


def test_shard_bounds():
    from amici_b200.sharding import shard_bounds
    for n in (0, 1, 7, 65536, 65537):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_sharded_objective_two_ranks(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "steadystate.npz"))
    total_llh, total_sllh = g["truth_llh"].sum(), g["truth_sllh"].sum(0)
    for rank, lo, hi, obj in res:
        np.testing.assert_allclose(obj[0], total_llh, rtol=1e-6)
        np.testing.assert_allclose(obj[1:-1], total_sllh, rtol=1e-4, atol=1e-6)
        assert obj[-1] == 0
    assert sorted((lo, hi) for _, lo, hi, _ in res) == [(0, 3), (3, 6)]


# ---- sharded objective with a parameter mapping (amici_b200/objective.py; reference simulate_petab / aggregate_sllh) ----
def _steadystate_problem():
    g = np.load(os.path.join(ROOT, "tests", "golden", "steadystate.npz"))
    N, npar = g["p"].shape
    index = np.tile(np.arange(npar), (N, 1))        # parameters 0..3 shared by all conditions
    index[:, 4] = -1                                 # parameter 4: condition specific ...
    index[0, 4], index[1, 4] = 4, 5                  # ... estimated for conditions 0 and 1, a fixed number elsewhere
    value = np.zeros((N, npar))
    value[:, 4] = g["p"][:, 4]
    theta = np.concatenate([g["p"][0, :4], [g["p"][0, 4], g["p"][1, 4]], [123.0]])   # last entry is mapped nowhere
    return g, index, value, theta


def _emul_solve(g, lo, hi):
    """Per-shard solve through the host build of the solver source (test aid standing in for the CUDA engine)."""
    import torch
    from amici_b200 import _capi
    lib = C.CDLL(os.path.join(ROOT, "tests", "cpu_emul", "libemul.so"))
    dims = dict(nx=3, np=5, nk=4, ny=3)

    def solve(p):
        b, out, keep = _capi.make_batch(dims, g["ts"], p.numpy(), g["k"][lo:hi], my=g["my"][lo:hi], sigmay=g["sigmay_in"][lo:hi])
        o = _capi.default_options(rtol=1e-8, atol=1e-12, sensi=1)
        assert lib.emul_run(b"steadystate", C.byref(o), C.byref(b)) == 0
        return torch.as_tensor(out["status"]), torch.as_tensor(out["llh"]), torch.as_tensor(out["sllh"])
    return solve


def _objective_worker(rank, world, port, q):
    import torch.distributed as dist
    from amici_b200.objective import ParameterMapping, ShardedObjective, shard_mapping
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, index, value, theta = _steadystate_problem()
    lo, hi, mp_ = shard_mapping(ParameterMapping(index, value), world, rank)
    obj = ShardedObjective(len(theta), mp_, np.arange(5), _emul_solve(g, lo, hi))
    llh, grad = obj(theta)
    q.put((rank, llh, grad))
    dist.destroy_process_group()


def test_objective_parameter_mapping_two_ranks(built):
    """fill -> solve -> aggregate -> one all-reduce gives, on every rank, what the reference's simulate_petab computes
    serially: llh = sum over conditions, sllh accumulated per problem parameter through the mapping (aggregate_sllh)."""
    import torch.multiprocessing as mp
    from amici_b200.objective import ParameterMapping, ShardedObjective
    g, index, value, theta = _steadystate_problem()
    N = len(index)
    serial = ShardedObjective(len(theta), ParameterMapping(index, value), np.arange(5), _emul_solve(g, 0, N))
    llh1, grad1 = serial(theta)
    status, llh_n, sllh_n = (t.numpy() for t in serial.last)
    assert (status == 0).all()
    # the reference's accumulation loop (aggregate_sllh, _simulations.py:313-357), written out
    acc = {}
    for n in range(N):
        for j in range(5):
            key = int(index[n, j])
            if key >= 0:
                acc[key] = acc.get(key, 0.0) + sllh_n[n, j]
    np.testing.assert_allclose(llh1, llh_n.sum(), rtol=1e-14)
    for key in range(len(theta)):
        np.testing.assert_allclose(grad1[key], acc.get(key, 0.0), rtol=1e-13, atol=1e-300)
    assert grad1[6] == 0.0 and set(acc) == {0, 1, 2, 3, 4, 5}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_objective_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, llh, grad in res:
        np.testing.assert_allclose(llh, llh1, rtol=1e-13)
        np.testing.assert_allclose(grad, grad1, rtol=1e-12, atol=1e-14)


def test_objective_failed_condition_gives_nan_and_no_gradient():
    import torch
    from amici_b200.objective import ParameterMapping, ShardedObjective
    index = np.array([[0, 1], [0, -1]])

    def solve(p):
        return torch.tensor([0, -4], dtype=torch.int32), torch.tensor([1.0, float("nan")]), torch.zeros(2, 2, dtype=torch.float64)
    llh, grad = ShardedObjective(2, ParameterMapping(index), [0, 1], solve)(np.array([1.0, 2.0]))
    assert np.isnan(llh) and grad is None


def test_objective_fills_fixed_parameters_from_problem_parameters():
    """map_sim_fix / map_preeq_fix: k and k_preeq get the UNSCALED problem parameter (or the condition table's number), the
    free parameters stay on their scale; fixed parameters add nothing to the gradient."""
    import torch
    from amici_b200.objective import FixedParameterMapping, ParameterMapping, ShardedObjective
    seen = {}

    def solve(p, k, k_preeq):
        seen.update(p=p.clone(), k=k.clone(), k_preeq=k_preeq)
        return torch.zeros(2, dtype=torch.int32), torch.tensor([1.0, 2.0], dtype=torch.float64), torch.ones(2, 1, dtype=torch.float64)
    theta = np.array([-1.0, 2.0, np.log(3.0)])
    obj = ShardedObjective(3, ParameterMapping(np.array([[0], [0]])), [0], solve,
                           fixed=FixedParameterMapping(np.array([[1, -1], [2, 0]]), np.array([[0.0, 7.0], [0.0, 0.0]])),
                           theta_scale=[2, 0, 1])
    llh, grad = obj(theta)
    np.testing.assert_allclose(seen["p"].numpy(), [[-1.0], [-1.0]])
    np.testing.assert_allclose(seen["k"].numpy(), [[2.0, 7.0], [3.0, 0.1]], rtol=1e-15)
    assert seen["k_preeq"] is None and llh == 3.0
    np.testing.assert_array_equal(grad, [2.0, 0.0, 0.0])
    with pytest.raises(ValueError):
        ShardedObjective(3, ParameterMapping(np.array([[0], [0]])), [0], solve, fixed=FixedParameterMapping(np.array([[3, -1], [2, 0]])))
