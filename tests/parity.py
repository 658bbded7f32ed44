"""Parity protocol shared by the CPU (lane-group emulation) and GPU tests.

The reference does not agree with itself to 10*rtol / 10*atol on sensitivities at rtol=1e-8, atol=1e-12
(KLU vs dense linear solver, simultaneous vs staggered), because step sequences are
chaotic w.r.t. round-off.  Every fixture therefore stores three reference runs: `klu` (the default
path), `dense` (same tolerances, other linear solver) and `truth` (rtol 1e-12 / atol 1e-16).

A candidate passes for a field iff
  (a) discrete outputs (status) are identical,
  (b) its error against `truth` is at most FACTOR x the reference's own error against `truth`
      (max of klu, dense), measured in the scaled norm |a-b| / (atol_f + rtol_f*|truth|), and
  (c) it deviates from `klu` by at most FACTOR x max(klu-vs-dense deviation, reference error against truth)
      in the same norm (two approximations of the truth may differ by the sum of their errors), or by
      less than 1 in that norm (i.e. inside 10*rtol / 10*atol outright).
"""
import numpy as np

RTOL, ATOL = 1e-8, 1e-12          # solver tolerances of the fixtures
FACTOR = 4.0
FLOAT_FIELDS = ["x", "y", "llh", "chi2", "sigmay", "res", "x0", "xdot", "J", "sx", "sy", "sllh", "ssigmay", "sres", "FIM", "sx0",
                "x_ss", "sx_ss", "z", "sz", "rz", "srz", "sigmaz"]
# discrete outputs that must be identical to the reference's
EXACT_FIELDS = ["status", "preeq_status"]


def scaled_err(a, b, truth, rtol=10 * RTOL, atol=10 * ATOL):
    a, b, truth = np.asarray(a, float), np.asarray(b, float), np.asarray(truth, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert (nan_a == nan_b).all(), "NaN pattern differs"
    err = np.abs(a - b) / (atol + rtol * np.abs(truth))
    return float(np.nanmax(err)) if err.size and not nan_a.all() else 0.0


def _propagated_scale(f, gold, truth):
    """Magnitude that the relative tolerance refers to.  For every field it is |truth|, except for the residuals and chi2: they
    are differences of near-equal numbers, res = (y - my) / sigma, so a deviation of y INSIDE its tolerance moves res by
    delta_y / sigma, which relative to |res| is larger by |y| / |y - my| (~100 on the fixtures).  Their tolerance is the one of y
    carried through the formula: |y| / sigma for res, sum 2 |res| |y| / sigma for chi2 (never smaller than |truth|)."""
    truth = np.asarray(truth, float)
    if f not in ("res", "chi2") or "truth_y" not in gold or "truth_sigmay" not in gold or "truth_res" not in gold:
        return truth
    y, sig, res = (np.asarray(gold["truth_" + k_], float) for k_ in ("y", "sigmay", "res"))
    if y.size != res.size:
        return truth
    with np.errstate(invalid="ignore", divide="ignore"):
        sc = np.nan_to_num(np.abs(y) / sig, nan=0.0, posinf=0.0).reshape(res.shape)
    if f == "res":
        return np.maximum(np.abs(truth), sc)
    return np.maximum(np.abs(truth), (2.0 * np.abs(np.nan_to_num(res)) * sc).reshape(res.shape[0], -1).sum(-1).reshape(truth.shape))


def check_against_golden(cand: dict, gold, fields=None, factor=FACTOR, verbose=True):
    """cand: field -> array; gold: loaded npz. Returns report dict; raises AssertionError on failure."""
    report = {}
    for f in EXACT_FIELDS:
        if f in cand and f"klu_{f}" in gold:
            np.testing.assert_array_equal(cand[f], gold[f"klu_{f}"], err_msg=f"{f} differs")
    for f in fields or FLOAT_FIELDS:
        if f not in cand or f"klu_{f}" not in gold:
            continue
        klu, dense, truth = gold[f"klu_{f}"], gold[f"dense_{f}"], gold[f"truth_{f}"]
        scale_ref = _propagated_scale(f, gold, truth)
        others = [dense] + [gold[k_] for k_ in (f"s{j}_{f}" for j in range(64)) if k_ in gold]   # reference_triple(spread_runs=...)
        floor_truth = max(scaled_err(r, truth, scale_ref) for r in [klu] + others)
        floor_pair = max(scaled_err(r, klu, scale_ref) for r in others)
        e_truth = scaled_err(cand[f], truth, scale_ref)
        e_klu = scaled_err(cand[f], klu, scale_ref)
        report[f] = dict(err_vs_truth=e_truth, ref_err_vs_truth=floor_truth, dev_vs_klu=e_klu, ref_dev=floor_pair)
        if verbose:
            print(f"{f:8s} cand-truth {e_truth:10.3e} (ref {floor_truth:10.3e})   cand-klu {e_klu:10.3e} (dense-klu {floor_pair:10.3e})")
        assert e_truth <= max(1.0, factor * floor_truth), (f, report[f])
        assert e_klu <= max(1.0, factor * max(floor_pair, floor_truth)), (f, report[f])
    return report


def robertson_view(g, sensi):
    """tests/golden/robertson.npz holds a plain forward run (`*_nosens_*`) and a forward-sensitivity run; select one."""
    v = {}
    for name in ("klu", "dense", "truth"):
        for f in ("status", "x", "y", "numsteps"):
            if f"{name}_{f}" in g:
                v[f"{name}_{f}"] = g[f"{name}_{f}"] if sensi else g[f"{name}_nosens_{f}"]
        if sensi:
            v[f"{name}_sx"], v[f"{name}_sy"] = g[f"{name}_sx"], g[f"{name}_sy"]
    return v


def reference_triple(ref, ts, p, k, fields, num_threads=1, truth=True, spread_runs=0, **kw):
    """Three runs of the compiled reference on the same inputs — KLU, dense, tight-tolerance truth — in the layout of a
    golden fixture, so that a fresh batch is checked by the same protocol as the committed vectors (check_against_golden).
    spread_runs > 0 adds that many runs `s0`, `s1`, ... with immaterially perturbed tolerances (rtol x (1 + 1e-6 (j + 1)),
    alternating linear solver): step sequences are chaotic w.r.t. such perturbations, so these sample the reference's own
    error distribution — two runs on a handful of conditions underestimate its tail (sx error of 512 steadystate conditions:
    median 2.4, p99 35, max 60 tolerances)."""
    out = {}
    truth_tol = kw.pop("truth_tol", (1e-12, 1e-16))   # model_calvetti_py: the reference's IDAS fails below 1e-10 / 1e-14
    runs = [("klu", dict(linsol=9)), ("dense", dict(linsol=1))]
    base_rtol = kw.get("rtol", RTOL)
    for j in range(spread_runs):
        runs.append((f"s{j}", dict(linsol=9 if j % 2 else 1, rtol=base_rtol * (1.0 + 1e-6 * (j + 1)))))
    if truth:
        runs.append(("truth", dict(linsol=9, rtol=truth_tol[0], atol=truth_tol[1], maxsteps=500000)))
    for name, extra in runs:
        r = ref.run(ts, p, k, fields=list(fields), num_threads=num_threads, **{**kw, **extra})
        for f in r:
            out[f"{name}_{f}"] = r[f]
    return out


def per_condition_err(a, b, truth, rtol=10 * RTOL, atol=10 * ATOL):
    """max over the elements of one condition of |a - b| / (10 atol + 10 rtol |truth|); arrays are [N, ...]."""
    a, b, truth = (np.asarray(v, float) for v in (a, b, truth))
    b, truth = b.reshape(a.shape), truth.reshape(a.shape)
    n = a.shape[0]
    err = np.abs(a - b) / (atol + rtol * np.abs(truth))
    err = np.where(np.isnan(a) & np.isnan(b), 0.0, err)     # same NaN (an unset z slot): no error
    return err.reshape(n, -1).max(axis=1) if err.size else np.zeros(n)


QUANTILES = (50, 90, 99)


def check_distribution(cand: dict, gold: dict, fields, ok=None, factor=FACTOR, qfactor=1.5, verbose=True):
    """Parity protocol for LARGE batches (hundreds to thousands of reference-checked conditions).  A maximum over thousands of
    chaotic step sequences is an extreme-value statistic, so next to the max-based rule of check_against_golden this
    compares DISTRIBUTIONS over conditions of the per-condition scaled error:

      accuracy  : every quantile (50 / 90 / 99 %) of the candidate's error against `truth` is at most qfactor x the same
                  quantile of the reference's own error (max of klu, dense), or inside the stated tolerance outright (<= 1);
      deviation : every quantile of the candidate's deviation from `klu` is at most factor x max(quantile of dense-vs-klu,
                  quantile of the reference's error against truth), or <= 1;
      maximum   : the rule of check_against_golden on the maxima.

    `ok`: boolean mask of the conditions to compare (those that succeed in every run)."""
    report = {}
    for f in fields:
        c = np.asarray(cand[f], float)
        klu, dense, truth = (np.asarray(gold[f"{n}_{f}"], float).reshape(c.shape) for n in ("klu", "dense", "truth"))
        if ok is not None:
            c, klu, dense, truth = c[ok], klu[ok], dense[ok], truth[ok]
        assert (np.isnan(c) == np.isnan(klu)).all(), f"{f}: NaN pattern differs from the reference's"
        e_c, e_k, e_d = (per_condition_err(v, truth, truth) for v in (c, klu, dense))
        d_c, d_d = per_condition_err(c, klu, truth), per_condition_err(dense, klu, truth)
        row = {}
        for q in QUANTILES:
            qc, qr = np.percentile(e_c, q), max(np.percentile(e_k, q), np.percentile(e_d, q))
            dc, dr = np.percentile(d_c, q), max(np.percentile(d_d, q), qr)
            row[q] = (qc, qr, dc, np.percentile(d_d, q))
            assert qc <= max(1.0, qfactor * qr), (f, "accuracy", q, qc, qr)
            assert dc <= max(1.0, factor * dr), (f, "deviation", q, dc, dr)
        mc, mr = e_c.max(), max(e_k.max(), e_d.max())
        md, mdr = d_c.max(), max(d_d.max(), mr)
        row["max"] = (mc, mr, md, d_d.max())
        assert mc <= max(1.0, factor * mr), (f, "max accuracy", mc, mr)
        assert md <= max(1.0, factor * mdr), (f, "max deviation", md, mdr)
        report[f] = row
        if verbose:
            for q, (qc, qr, dc, dd) in row.items():
                print(f"{f:6s} {str(q):>4s}  cand-truth {qc:10.3e} (ref {qr:10.3e})   cand-klu {dc:10.3e} (dense-klu {dd:10.3e})")
    return report
