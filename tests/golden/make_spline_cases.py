#!/usr/bin/env python
"""Writes tests/golden/spline_cases.npz — known answers for the Hermite spline (VERDICT r01 task 1e).

Two sources, both only readable in the build container (they do not travel; the fixture does):

1. The LITERAL tables of the reference's own unit tests, reference tests/cpp/unittests/testSplines.cpp:66-720, parsed here
   from the C++ source: for every TEST the constructor arguments, dvaluesdp / dslopesdp, the (t, value) or (t, sensitivities)
   expectations with the tolerances the test passes (ASSERT_DOUBLE_EQ = 4 ulp otherwise), the times at which it expects an
   exception, and the final value / final sensitivities it asserts.
2. Random splines over every valid combination of boundary condition x extrapolation x (finite-difference slopes) x
   (equidistant) x (logarithmic), evaluated by the UNMODIFIED reference class amici::HermiteSpline through
   oracle/_ref/libref_spline.so (oracle/ref_spline.cpp), incl. times before / after the node range, on nodes and t = inf.

    python tests/golden/make_spline_cases.py      (needs /root/reference and oracle/_ref)
"""
import ctypes as C
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AMICI_REFERENCE", "/root/reference")
NN, NP = 8, 4
BC = {"given": -1, "zeroDerivative": 0, "natural": 1, "naturalZeroDerivative": 2, "periodic": 3}
EP = {"noExtrapolation": -1, "constant": 0, "linear": 1, "polynomial": 2, "periodic": 3}


def _ev(txt):
    txt = re.sub(r"//[^\n]*", "", txt)
    txt = txt.replace("{", "[").replace("}", "]").replace("INFINITY", "float('inf')").replace("exp(", "np.exp(").replace("std::", "")
    return eval(txt, {"np": np})


def _top_level_split(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        if ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def parse_reference_tests():
    src = open(os.path.join(REF, "tests/cpp/unittests/testSplines.cpp")).read()
    blocks = re.split(r"\nTEST\(Splines, (\w+)\) \{", src)[1:]
    cases = []
    for name, body in zip(blocks[::2], blocks[1::2]):
        m = re.search(r"HermiteSpline spline\((.*?)\n    \);", body, re.S)
        args = _top_level_split(re.sub(r"//[^\n]*", "", m.group(1)))
        nodes, values, slopes = _ev(args[0]), _ev(args[1]), _ev(args[2])
        bc = [BC[a.split("::")[1]] for a in args[3:5]]
        ep = [EP[a.split("::")[1]] for a in args[5:7]]
        fd, equi, logp = [a.strip() == "true" for a in args[7:10]]
        case = dict(name=name, nodes=nodes, values=values, slopes=slopes, bc=bc, ep=ep, fd=fd, equi=equi, log=logp, npar=0, dv=[], dm=[],
                    tv=[], ts_=[], throws=[], final=None, rtol=0.0, atol=0.0, line=src[:src.index("TEST(Splines, " + name)].count("\n") + 1)
        m = re.search(r"dvaluesdp\s*=\s*(\{.*?\});", body, re.S)
        if m:
            case["dv"] = _ev(m.group(1))
            case["npar"] = int(re.search(r"int n_params = (\d+);", body).group(1))
            m2 = re.search(r"std::vector<double> dslopesdp = (\{.*?\});", body, re.S)
            case["dm"] = _ev(m2.group(1)) if m2 else [0.0] * len(case["dv"])
        m = re.search(r"expectations\s*=\s*(\{.*?\n    \});", body, re.S)
        if m:
            table = _ev(m.group(1))
            if isinstance(table[0][1], list):
                case["ts_"] = table
            else:
                case["tv"] = table
            m = re.search(r"test_spline_(?:values|sensitivities)\(\s*spline, expectations(?:,\s*([^,\)]+),\s*([^\)]+))?\s*\)", body)
            if m.group(1):
                case["rtol"], case["atol"] = float(m.group(1)), float(m.group(2))
        case["throws"] = [float(_ev(t)) for t in re.findall(r"ASSERT_THROW\(spline\.get_(?:value|sensitivity)\(([^,\)]+)", body)]
        fin = {}
        m = re.search(r"ASSERT_DOUBLE_EQ\(spline\.get_final_value\(\), ([^\)]+(?:\))?)\);", body)
        if m:
            fin["value"] = float(_ev(m.group(1)))
        elif "isnan(spline.get_final_value())" in body:
            fin["value"] = float("nan")
        for ip, val in re.findall(r"ASSERT_DOUBLE_EQ\(\s*spline\.get_final_sensitivity\((\d)\), ([^;]+)\);", body):
            fin.setdefault("sens", {})[int(ip)] = float(_ev(val))
        for ip in re.findall(r"isnan\(spline\.get_final_sensitivity\((\d)\)\)", body):
            fin.setdefault("sens", {})[int(ip)] = float("nan")
        case["final"] = fin or None
        cases.append(case)
    return cases


def pack(cases):
    """cases -> arrays in the layout of ab_selftest_spline (include/amici_b200.h)"""
    n = len(cases)
    cfg = np.zeros((n, 8), np.int32)
    nodes, values, slopes = (np.zeros((n, NN)) for _ in range(3))
    dv, dm = np.zeros((n, NN, NP)), np.zeros((n, NN, NP))
    for i, c in enumerate(cases):
        nn = len(c["values"])
        cfg[i] = [nn, c["bc"][0], c["bc"][1], c["ep"][0], c["ep"][1], c["fd"], c["equi"], c["log"]]
        nodes[i, :len(c["nodes"])] = c["nodes"]
        values[i, :nn] = c["values"]
        values[i, nn:] = 1.0
        if not c["fd"]:
            slopes[i, :nn] = c["slopes"]
        if c["npar"]:
            dv[i, :nn, :c["npar"]] = np.asarray(c["dv"], float).reshape(nn, c["npar"])
            dm[i, :nn, :c["npar"]] = np.asarray(c["dm"], float).reshape(nn, c["npar"])
    return cfg, nodes, values, slopes, dv, dm


def reference_eval(cfg, nodes, values, slopes, dv, dm, ts):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_spline.so"))
    dp, ip_ = C.POINTER(C.c_double), C.POINTER(C.c_int)
    n, nt = ts.shape
    val, sens, thrown, rc = np.zeros((n, nt)), np.zeros((n, nt, NP)), np.zeros((n, nt), np.int32), np.zeros(n, np.int32)
    for i in range(n):
        nn = int(cfg[i, 0])
        x = np.ascontiguousarray(nodes[i, :2] if cfg[i, 6] else nodes[i, :nn])
        v, m = np.ascontiguousarray(values[i, :nn]), np.ascontiguousarray(slopes[i, :nn])
        dvi, dmi = np.ascontiguousarray(dv[i, :nn, :]), np.ascontiguousarray(dm[i, :nn, :])
        t = np.ascontiguousarray(ts[i])
        vo, so, th = np.zeros(nt), np.zeros((nt, NP)), np.zeros(nt, np.int32)
        rc[i] = lib.ref_spline_eval(nn, len(x), x.ctypes.data_as(dp), v.ctypes.data_as(dp), m.ctypes.data_as(dp), int(cfg[i, 1]), int(cfg[i, 2]),
                                    int(cfg[i, 3]), int(cfg[i, 4]), int(cfg[i, 5]), int(cfg[i, 6]), int(cfg[i, 7]), NP, dvi.ctypes.data_as(dp),
                                    dmi.ctypes.data_as(dp), nt, t.ctypes.data_as(dp), vo.ctypes.data_as(dp), so.ctypes.data_as(dp), th.ctypes.data_as(ip_))
        val[i], sens[i], thrown[i] = vo, so, th
    return val, sens, thrown, rc


def random_cases(rng, per_combo=1):
    cases = []
    for bcf in (-1, 0, 1, 3):
        for bcl in (-1, 0, 1, 3):
            if (bcf == 3) != (bcl == 3):
                continue
            for epf in (-1, 0, 1, 2, 3):
                for epl in (-1, 0, 1, 2, 3):
                    for fd in (True, False):
                        for equi in (True, False):
                            for logp in (True, False):
                                if fd and 1 in (bcf, bcl):
                                    continue       # natural + finite differences: no sensitivities in the reference
                                if (epf == 1 and bcf not in (-1, 0)) or (epl == 1 and bcl not in (-1, 0)):
                                    continue       # linear extrapolation sensitivities exist for given / zeroDerivative only
                                for _ in range(per_combo):
                                    nn = int(rng.integers(3, NN + 1))
                                    x = np.sort(rng.uniform(-1, 4, nn)) if not equi else np.array([rng.uniform(-1, 0.5), rng.uniform(1, 4)])
                                    if not equi and np.min(np.diff(x)) < 0.05:
                                        x = np.linspace(x[0], x[0] + 3, nn) + rng.uniform(-0.02, 0.02, nn)
                                        x.sort()
                                    v = rng.uniform(0.2, 3.0, nn)
                                    if bcf == 3:
                                        v[-1] = v[0]
                                    cases.append(dict(nodes=list(x), values=list(v), slopes=list(rng.normal(0, 1, nn)), bc=[bcf, bcl], ep=[epf, epl],
                                                      fd=fd, equi=equi, log=logp, npar=NP, dv=list(rng.normal(0, 1, nn * NP)),
                                                      dm=list(rng.normal(0, 1, nn * NP))))
    return cases


def main():
    ref_cases = parse_reference_tests()
    print(f"parsed {len(ref_cases)} TESTs of testSplines.cpp")
    out = {}
    # ---- 1. literal tables -----------------------------------------------------------------------------------------------------
    cfg, nodes, values, slopes, dv, dm = pack(ref_cases)
    NT = 16
    ts = np.zeros((len(ref_cases), NT))
    kind = np.zeros((len(ref_cases), NT), np.int32)      # 0 unused, 1 value row, 2 sensitivity row, 3 must throw, 4 final value / sensitivities
    expect_v = np.full((len(ref_cases), NT), np.nan)
    expect_s = np.full((len(ref_cases), NT, NP), np.nan)
    tol = np.zeros((len(ref_cases), 2))
    for i, c in enumerate(ref_cases):
        j = 0
        for t, v in c["tv"]:
            ts[i, j], kind[i, j], expect_v[i, j] = t, 1, v
            j += 1
        for t, s in c["ts_"]:
            ts[i, j], kind[i, j] = t, 2
            expect_s[i, j, :len(s)] = s
            j += 1
        for t in c["throws"]:
            ts[i, j], kind[i, j] = t, 3
            j += 1
        if c["final"]:
            ts[i, j], kind[i, j] = np.inf, 4
            expect_v[i, j] = c["final"].get("value", np.nan)
            for ip, s in c["final"].get("sens", {}).items():
                expect_s[i, j, ip] = s
            out_final_has_value = "value" in c["final"]
            kind[i, j] = 4 if out_final_has_value else 5
            j += 1
        assert j <= NT, (c["name"], j)
        ts[i, j:] = c["tv"][0][0] if c["tv"] else (c["ts_"][0][0] if c["ts_"] else 0.5 * (c["nodes"][0] + c["nodes"][-1]))
        tol[i] = c["rtol"], c["atol"]
    out.update(lit_cfg=cfg, lit_nodes=nodes, lit_values=values, lit_slopes=slopes, lit_dv=dv, lit_dm=dm, lit_ts=ts, lit_kind=kind,
               lit_expect_v=expect_v, lit_expect_s=expect_s, lit_tol=tol, lit_npar=np.array([c["npar"] for c in ref_cases], np.int32),
               lit_names=np.array([f"{c['name']} (testSplines.cpp:{c['line']})" for c in ref_cases]))
    # the compiled reference must reproduce its own tables (pins the oracle wrapper)
    val, sens, thrown, rc = reference_eval(cfg, nodes, values, slopes, dv, dm, ts)
    assert (rc == 0).all()
    out.update(lit_ref_v=val, lit_ref_s=sens, lit_ref_thrown=thrown)
    # ---- 2. random configurations through the live reference -------------------------------------------------------------------
    rng = np.random.default_rng(20261018)
    rc_cases = random_cases(rng)
    cfg, nodes, values, slopes, dv, dm = pack(rc_cases)
    n = len(rc_cases)
    NT2 = 24
    ts = np.zeros((n, NT2))
    for i, c in enumerate(rc_cases):
        nn = len(c["values"])
        x = np.asarray(c["nodes"]) if not c["equi"] else np.linspace(c["nodes"][0], c["nodes"][1], nn)
        span = x[-1] - x[0]
        inside = rng.uniform(x[0], x[-1], 10)
        outside = np.concatenate([x[0] - rng.uniform(0.01, 2.5, 4) * span, x[-1] + rng.uniform(0.01, 2.5, 4) * span])
        ts[i] = np.concatenate([inside, outside, [x[0], x[-1], x[1], x[nn // 2], x[nn - 2]], [np.inf]])
    val, sens, thrown, rc = reference_eval(cfg, nodes, values, slopes, dv, dm, ts)
    keep = rc == 0
    print(f"{n} random configurations, {int(keep.sum())} accepted by the reference's constructor")
    out.update(rnd_cfg=cfg[keep], rnd_nodes=nodes[keep], rnd_values=values[keep], rnd_slopes=slopes[keep], rnd_dv=dv[keep], rnd_dm=dm[keep],
               rnd_ts=ts[keep], rnd_ref_v=val[keep], rnd_ref_s=sens[keep], rnd_ref_thrown=thrown[keep])
    np.savez_compressed(os.path.join(HERE, "spline_cases.npz"), **out)
    print("wrote", os.path.join(HERE, "spline_cases.npz"))


if __name__ == "__main__":
    sys.exit(main())
