"""Parity at BASELINE sizes (VERDICT r01 task 1b): every config of BASELINE.json is integrated at its FULL size on the GPU
through the C ABI (65,536 conditions; 4096 for Robertson), and the first K conditions of the same seeded batch are integrated
by the compiled, unmodified reference on the host cores three times — default linear solver (KLU), dense, tight-tolerance
truth (rtol 1e-12 / atol 1e-16).  Assertions:

  * discrete outcomes identical: status, per-root event counts, SteadyStateStatus triples, Newton step counts;
  * step-count statistics of the whole sample within 0.5 %;
  * floats by the parity protocol on DISTRIBUTIONS over conditions (tests/parity.py::check_distribution): the engine's error
    against the truth is the reference's own error at the 50 / 90 / 99 % quantiles and at the maximum, its deviation from
    the KLU run is at the reference's own KLU-vs-dense spread;
  * size-independent properties on all 65,536 conditions (finite outputs, status distribution, conservation laws).

Inputs are bench.py's workloads, so what is timed is what is checked."""
import os
import sys

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _engine_and_reference(name, K, fields, extra_ref_fields=()):
    import amici_b200 as ab
    import bench
    from oracle import refsim
    c = bench.CONFIGS[name]
    if not refsim.available(c["ref"]):
        pytest.skip("oracle/_ref not built")
    model, make_solver, w = bench.build_engine_objects(ab, c, c["n"], 0)
    ref = refsim.RefModel(c["ref"])
    sl = slice(0, K)
    my = sig = None
    if w["noise"] is not None:
        # measurements: noise-free observables of the reference on the checked slice, of the engine elsewhere (bench.py does the same)
        y = ab.run_batch(model, make_solver(False), ab.ExpDataBatch(c["ts"], w["p"], w["k"], k_preeq=w.get("k_preeq")), fields=["y"])["y"]
        my = y + w["noise"]
        sig = np.full_like(my, 0.1)
    batch = ab.ExpDataBatch(c["ts"], w["p"], w["k"], my=my, sigmay=sig, k_preeq=w.get("k_preeq"))
    want = ["status", "numsteps"] + list(fields) + list(extra_ref_fields)
    res = ab.run_batch(model, make_solver(c["sensi"]), batch, fields=want)
    kw = {k_: v for k_, v in bench.reference_kwargs(c).items() if k_ not in ("rtol", "atol")}
    common = dict(k_preeq=None if "k_preeq" not in w else w["k_preeq"][sl], my=None if my is None else my[sl],
                  sigmay=None if sig is None else sig[sl])
    gold = parity.reference_triple(ref, c["ts"], w["p"][sl], w["k"][sl], want, num_threads=ref.max_threads(), **common, **kw)
    return c, res, gold, sl


def _common_checks(res, gold, sl, fields, ok_fraction=1.0):
    st = res["status"][sl]
    np.testing.assert_array_equal(st, gold["klu_status"], err_msg="status codes differ from the reference's")
    ok = (st == 0) & (gold["dense_status"] == 0) & (gold["truth_status"] == 0)
    assert ok.mean() >= ok_fraction
    ns_e, ns_r = res["numsteps"][sl][ok][:, -1].astype(float), gold["klu_numsteps"][ok][:, -1].astype(float)
    assert abs(ns_e.mean() / ns_r.mean() - 1.0) < 5e-3, (ns_e.mean(), ns_r.mean())
    parity.check_distribution({f: np.asarray(res[f])[sl] for f in fields}, gold, fields, ok=ok)
    return ok


def test_baseline_config1_jakstat_65536():
    fields = ["x", "y", "llh", "sx", "sy", "sllh", "res", "sres", "FIM"]
    c, res, gold, sl = _engine_and_reference("jakstat", 2048, fields, extra_ref_fields=["chi2"])
    assert (res["status"] == 0).all()
    _common_checks(res, gold, sl, fields)
    for f in fields:
        assert np.isfinite(np.asarray(res[f])).all(), f
    # chi2 / res / llh are tied together exactly by the Gaussian noise model (reference src/rdata.cpp:1002-1060)
    np.testing.assert_allclose((np.asarray(res["res"]) ** 2).reshape(len(res), -1).sum(-1), np.asarray(res["chi2"]), rtol=1e-12)


def test_baseline_config0_robertson_4096_against_cvodes():
    """config 1 of SURVEY §8d: the ODE form through the reference's own CVODES (oracle/gen_ref_model.py writes the model in the
    reference's generated-code format), all 4096 draws checked."""
    fields = ["x", "y"]
    c, res, gold, sl = _engine_and_reference("robertson", 4096, fields)
    assert (res["status"] == 0).all()
    _common_checks(res, gold, sl, fields)
    x = np.asarray(res["x"])
    np.testing.assert_allclose(x.sum(-1), 1.0, rtol=0, atol=1e-9)


def test_baseline_config2_steadystate_65536():
    fields = ["x", "sx", "llh", "sllh", "x_ss", "sx_ss", "x0", "sx0"]
    c, res, gold, sl = _engine_and_reference("steadystate", 2048, fields, extra_ref_fields=["preeq_status", "preeq_numsteps"])
    assert (res["status"] == 0).all()
    np.testing.assert_array_equal(res["preeq_status"][sl], gold["klu_preeq_status"])
    np.testing.assert_array_equal(res["preeq_numsteps"][sl][:, [0, 2]], gold["klu_preeq_numsteps"][:, [0, 2]])
    sim = gold["klu_preeq_numsteps"][:, 1].astype(float)
    assert (np.abs(res["preeq_numsteps"][sl][:, 1] - sim) <= 0.15 * sim + 2).all()
    assert (res["preeq_status"][:, 0] == 1).mean() > 0.9     # Newton finds almost all steady states
    _common_checks(res, gold, sl, fields)


@pytest.mark.parametrize("name,K,okf", [("events", 2048, 0.99), ("neuron", 1024, 0.9)])
def test_baseline_config3_events_65536(name, K, okf):
    fields = ["x", "y", "sx", "z", "sz"]
    c, res, gold, sl = _engine_and_reference(name, K, fields, extra_ref_fields=["rz"])
    st_e, st_r = res["status"][sl], gold["klu_status"]
    # neuron: ~5 % of the draws spike so often that maxsteps runs out, in the reference too.  The same conditions fail; a failing one
    # runs out of steps inside CVode (-1) or on AMICI's own check (-99) depending on where the 10000th step lands
    np.testing.assert_array_equal(st_e == 0, st_r == 0)
    assert np.isin(st_e, (0, -1, -99)).all() and np.isin(st_r, (0, -1, -99)).all()
    ok = (st_e == 0) & (gold["dense_status"] == 0) & (gold["truth_status"] == 0)
    assert ok.mean() >= okf
    # identical event counts per root: an unset slot of z is NaN, a recorded event finite
    tmax = c["ts"][-1]
    z_e, z_r = np.asarray(res["z"])[sl][ok], gold["klu_z"][ok]
    np.testing.assert_array_equal(np.isfinite(z_e) & (z_e < tmax), np.isfinite(z_r) & (z_r < tmax), err_msg="event counts per root differ")
    ns_e, ns_r = res["numsteps"][sl][ok][:, -1].astype(float), gold["klu_numsteps"][ok][:, -1].astype(float)
    assert abs(ns_e.mean() / ns_r.mean() - 1.0) < 5e-3, (ns_e.mean(), ns_r.mean())
    parity.check_distribution({f: np.asarray(res[f])[sl] for f in fields}, gold, fields, ok=ok)
    good = res["status"] == 0
    assert good.mean() >= okf
    assert np.isfinite(np.asarray(res["sx"])[good]).all()
