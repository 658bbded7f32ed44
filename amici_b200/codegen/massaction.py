"""Synthetic mass-action reaction networks (BASELINE.json configs[4]: 100 species / 300 reactions).

The reference would obtain such a model from SBML through its importer (libsbml, not available in this image); here the
network is generated directly (seeded) as *tables*, and the same tables feed

* the CUDA header for the block solver (``emit_cuda``: ``csrc/models/<name>.cuh``, consumed by ``csrc/bdf_block.cuh``), and
* the reference-format C++ model sources the oracle compiles against the unmodified libamici
  (``oracle/gen_massaction_ref.py``).

Kinetics: reaction j has one or two reactants and one or two products, rate ``v_j = p_j * x_a [* x_b]``; the rate constants
are the free parameters (np = number of reactions).  Species carry a mass of 1 (monomers) or 2 (dimers) and every reaction
conserves mass (A->B, D->E, A+B->D, D->A+B), so trajectories stay bounded and non-negative for any positive rate constants.
"""
from __future__ import annotations

import os

import numpy as np

SPECS = {
    # name: (monomers, dimers, [A->B, D->E, A+B->D, D->A+B], n_observables, seed)
    "massaction": (60, 40, (90, 60, 75, 75), 10, 5),
    "massaction_small": (6, 4, (8, 4, 6, 6), 3, 6),
}


def network(name: str) -> dict:
    n1, n2, counts, ny, seed = SPECS[name]
    rng = np.random.default_rng(seed)
    nx, nr = n1 + n2, sum(counts)
    r1, r2, q1, q2 = [], [], [], []   # reactants (r2 = -1: none), products (q2 = -1: none)
    for _ in range(counts[0]):
        a, b = rng.choice(n1, 2, replace=False)
        r1.append(a); r2.append(-1); q1.append(b); q2.append(-1)
    for _ in range(counts[1]):
        a, b = n1 + rng.choice(n2, 2, replace=False)
        r1.append(a); r2.append(-1); q1.append(b); q2.append(-1)
    for _ in range(counts[2]):
        a, b = rng.choice(n1, 2, replace=False)
        r1.append(a); r2.append(b); q1.append(n1 + rng.integers(n2)); q2.append(-1)
    for _ in range(counts[3]):
        a, b = rng.choice(n1, 2, replace=False)
        r1.append(n1 + rng.integers(n2)); r2.append(-1); q1.append(a); q2.append(b)
    order = rng.permutation(nr)
    r1, r2, q1, q2 = (np.asarray(v, dtype=np.int64)[order] for v in (r1, r2, q1, q2))
    p_nominal = 10.0 ** rng.uniform(-0.5, 0.5, nr)
    x0 = rng.uniform(0.5, 1.5, nx)
    obs = np.sort(rng.choice(nx, ny, replace=False))
    net = dict(name=name, nx=nx, nr=nr, ny=ny, nk=1, r1=r1, r2=r2, q1=q1, q2=q2, p_nominal=p_nominal, k_nominal=np.array([1.0]),
               x0=x0, obs=obs, sigma=0.1)
    _derive(net)
    return net


def _derive(net: dict) -> None:
    """Stoichiometry lists and the static structure of J = d xdot / d x."""
    nx, nr = net["nx"], net["nr"]
    r1, r2, q1, q2 = net["r1"], net["r2"], net["q1"], net["q2"]
    # participants of reaction j: (species, stoichiometric coefficient)
    part = []
    for j in range(nr):
        pj = [(int(r1[j]), -1.0)]
        if r2[j] >= 0:
            pj.append((int(r2[j]), -1.0))
        pj.append((int(q1[j]), 1.0))
        if q2[j] >= 0:
            pj.append((int(q2[j]), 1.0))
        part.append(pj)
    net["part"] = part
    # species -> [(reaction, coefficient)]
    sp = [[] for _ in range(nx)]
    for j, pj in enumerate(part):
        for i, c in pj:
            sp[i].append((j, c))
    net["sp"] = sp
    # J[i][k] = sum over terms (reaction j, partner species or -1, coefficient): coef * p_j * (x_partner or 1)
    terms: dict = {}
    for j, pj in enumerate(part):
        reac = [int(r1[j])] + ([int(r2[j])] if r2[j] >= 0 else [])
        for idx, kx in enumerate(reac):
            partner = reac[1 - idx] if len(reac) == 2 else -1
            for i, c in pj:
                terms.setdefault((i, kx), []).append((j, partner, c))
    net["jterms"] = terms   # keyed by (row, col)


def _csr(net):
    """CSR (row-major) arrays of J with its term lists."""
    nx = net["nx"]
    keys = sorted(net["jterms"].keys())
    rowptr, col, tptr, trxn, tpart, tcoef = [0], [], [0], [], [], []
    r = 0
    for (i, k) in keys:
        while r < i:
            rowptr.append(len(col)); r += 1
        col.append(k)
        for (j, partner, c) in net["jterms"][(i, k)]:
            trxn.append(j); tpart.append(partner); tcoef.append(c)
        tptr.append(len(trxn))
    while r < nx:
        rowptr.append(len(col)); r += 1
    return rowptr, col, tptr, trxn, tpart, tcoef


def csc_pattern(net):
    """(colptr, rowval) of J in the reference's column-major convention, plus the (row, col) key per entry."""
    nx = net["nx"]
    keys = sorted(net["jterms"].keys(), key=lambda rc: (rc[1], rc[0]))
    colptr, rowval = [0], []
    c = 0
    for (i, k) in keys:
        while c < k:
            colptr.append(len(rowval)); c += 1
        rowval.append(i)
    while c < nx:
        colptr.append(len(rowval)); c += 1
    return colptr, rowval, keys


def emit_cuda(net: dict) -> str:
    name, nx, nr, ny = net["name"], net["nx"], net["nr"], net["ny"]
    rowptr, col, tptr, trxn, tpart, tcoef = _csr(net)
    spptr, sprxn, spcoef = [0], [], []
    for lst in net["sp"]:
        for j, c in lst:
            sprxn.append(j); spcoef.append(int(c))
        spptr.append(len(sprxn))
    ptptr, ptsp, ptcoef = [0], [], []
    for pj in net["part"]:
        for i, c in pj:
            ptsp.append(i); ptcoef.append(int(c))
        ptptr.append(len(ptsp))
    rc = []
    for i in range(nx):
        assert rowptr[i + 1] > rowptr[i], "every species must take part in a reaction"
        rc += [(i << 16) | k for k in col[rowptr[i]:rowptr[i + 1]]]
    guard = f"AMICI_B200_MODEL_{name.upper()}_CUH"
    tables = [   # (C type, name, values, comment)
        ("int", "r1", [int(v) for v in net["r1"]], "first reactant of reaction j"),
        ("int", "r2", [int(v) for v in net["r2"]], "second reactant of reaction j, -1 if none"),
        ("int", "sp_ptr", spptr, "species -> (reaction, stoichiometric coefficient): xdot_i = sum coef * v_j"),
        ("int", "sp_rxn", sprxn, ""), ("signed char", "sp_coef", spcoef, ""),
        ("int", "pt_ptr", ptptr, "reaction -> (species, stoichiometric coefficient): d xdot / d p_j = coef * v_j / p_j"),
        ("int", "pt_sp", ptsp, ""), ("signed char", "pt_coef", ptcoef, ""),
        ("int", "j_rowptr", rowptr, "J = d xdot / d x in CSR; entry e = sum over its terms of coef * p[rxn] * (partner >= 0 ? x[partner] : 1)"),
        ("int", "j_col", col, ""), ("int", "j_rc", rc, "row << 16 | column of entry e (flat sweep over the entries)"),
        ("int", "jt_ptr", tptr, ""), ("int", "jt_rxn", trxn, ""),
        ("int", "jt_partner", tpart, ""), ("signed char", "jt_coef", [int(c) for c in tcoef], ""),
        ("int", "obs", [int(v) for v in net["obs"]], "observable iy = x[obs[iy]]"),
        ("double", "x0", [repr(float(v)) for v in net["x0"]], "initial state (constants)"),
    ]
    L = []
    A = L.append
    A("// GENERATED by amici_b200/codegen/massaction.py — do not edit.")
    A(f"// Tables of the synthetic mass-action network '{name}' ({nx} species, {nr} reactions) for the block solver")
    A("// (csrc/bdf_block.cuh). Rate law v_j = p_j * x[r1_j] * (r2_j >= 0 ? x[r2_j] : 1); xdot = N v.")
    A(f"#ifndef {guard}\n#define {guard}\nnamespace amici_b200 {{ namespace models {{\n")
    A(f"namespace {name}_tab {{")
    for ctype, tname, vals, comment in tables:
        if comment:
            A(f"// {comment}")
        body = ", ".join(str(v) for v in vals) if len(vals) else "0"
        A(f"__device__ const {ctype} {tname}[{max(1, len(vals))}] = {{{body}}};")
    A(f"}}  // namespace {name}_tab\n")
    A(f"struct {name} {{")
    A(f"    static constexpr int NX = {nx}, NP = {nr}, NK = 1, NY = {ny}, NZ = 0, NE = 0, NSPL = 0;")
    A(f"    static constexpr int NNZ_J = {len(col)}, NTERMS = {len(trxn)};")
    A(f"    static constexpr double SIGMA = {net['sigma']!r};   // model noise level of every observable")
    A(f"    static const char* name() {{ return \"{name}\"; }}")
    for ctype, tname, vals, _ in tables:
        A(f"    static __device__ __forceinline__ {ctype} {tname}(const int i) {{ return {name}_tab::{tname}[i]; }}")
    A("    static void nominal(double* p, double* k) {")
    A("        static const double pn[] = {" + ", ".join(repr(float(v)) for v in net["p_nominal"]) + "};")
    A(f"        for (int i = 0; i < {nr}; ++i) p[i] = pn[i];")
    A("        k[0] = 1.0;")
    A("    }")
    A("};\n")
    A("}}  // namespace amici_b200::models")
    A(f"#endif  // {guard}")
    return "\n".join(L) + "\n"


def generate_all(out_dir: str | None = None) -> list[str]:
    out_dir = out_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "models")
    paths = []
    for name in SPECS:
        path = os.path.join(out_dir, f"{name}.cuh")
        txt = emit_cuda(network(name))
        if not os.path.exists(path) or open(path).read() != txt:
            with open(path, "w") as f:
                f.write(txt)
        paths.append(path)
    return paths


# ---- numpy restatement (host logic / tests): xdot, J, d xdot / d p -------------------------------------------------
def xdot(net, x, p):
    u = x[net["r1"]] * np.where(net["r2"] >= 0, x[np.maximum(net["r2"], 0)], 1.0)
    v = p * u
    out = np.zeros(net["nx"])
    for j, pj in enumerate(net["part"]):
        for i, c in pj:
            out[i] += c * v[j]
    return out


if __name__ == "__main__":
    for p in generate_all():
        print("wrote", p)
