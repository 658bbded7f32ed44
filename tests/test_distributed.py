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
