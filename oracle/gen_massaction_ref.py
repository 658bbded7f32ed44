#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE — writes a mass-action network (amici_b200/codegen/massaction.py) as model sources in the
format of the reference's own code generator, so that the UNMODIFIED libamici (CVODES + KLU) integrates the same model the
block solver integrates.

  python oracle/gen_massaction_ref.py <name> <out_dir> <reference root>

Writes <out_dir>/model_<name>_py/{*.cpp, *.h}.  The per-function files are written from the network tables with the
plain-pointer signatures of the reference's generated code (reference python/sdist/amici/exporters/sundials/
cxx_functions.py:101-418; a shipped example of the layout: models/model_steadystate_py/).  The model class header — pure
boilerplate wiring those functions into amici::Model_ODE — is derived at build time from the shipped header of a model with
the same function set (models/model_steadystate_py/model_steadystate_py.h, read from the reference tree where it lies) by
substituting the model name, the dimensions and the default parameters; nothing of it is stored in this repository.
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amici_b200.codegen.massaction import csc_pattern, network  # noqa: E402

PROLOGUE = '#include "amici/symbolic_functions.h"\n#include "amici/defines.h"\n#include "amici/sundials_matrix_wrapper.h"\n' \
           '#include "sundials/sundials_types.h"\n#include <gsl/gsl-lite.hpp>\n#include <algorithm>\n#include <array>\n#include <cmath>\n'
SIG_X = "const realtype t, const realtype *x, const realtype *p, const realtype *k, const realtype *h, const realtype *w"
SIG_Y = "const int iy, const realtype *p, const realtype *k, const realtype *y, const realtype *sigmay, const realtype *my"


def _rate(net, j, drop=None):
    """p_j * x_a * x_b with one factor dropped: 'p', or a species index."""
    f = []
    if drop != "p":
        f.append(f"p[{j}]")
    reac = [int(net["r1"][j])] + ([int(net["r2"][j])] if net["r2"][j] >= 0 else [])
    dropped = False
    for s in reac:
        if drop is not None and drop != "p" and s == drop and not dropped:
            dropped = True
            continue
        f.append(f"x[{s}]")
    return "*".join(f) if f else "1.0"


def _signed_sum(terms):
    out = ""
    for c, txt in terms:
        out += (" - " if c < 0 else " + ") + txt
    return out[3:] if out.startswith(" + ") else "-" + out[3:]


def _index_fn(ns, fn, model, kind, vals):
    name = f"{fn}_{kind}_{model}"
    setter = "set_indexptrs" if kind == "colptrs" else "set_indexvals"
    return (f"static constexpr std::array<sunindextype, {len(vals)}> {name}_ = {{\n    {', '.join(str(v) for v in vals)}\n}};\n"
            f"void {name}(SUNMatrixWrapper &m){{\n    m.{setter}(gsl::make_span({name}_));\n}}\n")


def write_model(name: str, out_root: str, ref_root: str) -> str:
    net = network(name)
    model = f"model_{name}_py"
    ns = f"model_{model}"
    d = os.path.join(out_root, model)
    os.makedirs(d, exist_ok=True)
    nx, nr, ny = net["nx"], net["nr"], net["ny"]

    def put(fname, body, wrap=True):
        txt = PROLOGUE + f"namespace amici {{\nnamespace {ns} {{\n{body}}} // namespace {ns}\n}} // namespace amici\n" if wrap else body
        path = os.path.join(d, fname)
        if not os.path.exists(path) or open(path).read() != txt:
            with open(path, "w") as f:
                f.write(txt)

    # xdot
    lines = []
    for i, lst in enumerate(net["sp"]):
        lines.append(f"    xdot[{i}] = " + (_signed_sum([(c, _rate(net, j)) for j, c in lst]) if lst else "0.0") + ";")
    put("xdot.cpp", f"void xdot_{model}(realtype *xdot, {SIG_X}){{\n" + "\n".join(lines) + "\n}\n")
    # dxdotdx_explicit (CSC)
    colptr, rowval, keys = csc_pattern(net)
    lines = []
    for e, (i, kx) in enumerate(keys):
        lines.append(f"    dxdotdx_explicit[{e}] = " + _signed_sum([(c, _rate(net, j, drop=kx)) for j, _, c in net["jterms"][(i, kx)]]) + ";")
    put("dxdotdx_explicit.cpp", _index_fn(ns, "dxdotdx_explicit", model, "colptrs", colptr) + _index_fn(ns, "dxdotdx_explicit", model, "rowvals", rowval) +
        f"void dxdotdx_explicit_{model}(realtype *dxdotdx_explicit, {SIG_X}){{\n" + "\n".join(lines) + "\n}\n")
    # dxdotdp_explicit (CSC over parameters)
    pcol, prow, lines = [0], [], []
    for j, pj in enumerate(net["part"]):
        for i, c in sorted(pj):
            lines.append(f"    dxdotdp_explicit[{len(prow)}] = " + _signed_sum([(c, _rate(net, j, drop='p'))]) + ";")
            prow.append(i)
        pcol.append(len(prow))
    put("dxdotdp_explicit.cpp", _index_fn(ns, "dxdotdp_explicit", model, "colptrs", pcol) + _index_fn(ns, "dxdotdp_explicit", model, "rowvals", prow) +
        f"void dxdotdp_explicit_{model}(realtype *dxdotdp_explicit, {SIG_X}){{\n" + "\n".join(lines) + "\n}\n")
    # observables and the Gaussian objective
    obs = [int(v) for v in net["obs"]]
    put("y.cpp", f"void y_{model}(realtype *y, {SIG_X}){{\n" + "\n".join(f"    y[{iy}] = x[{s}];" for iy, s in enumerate(obs)) + "\n}\n")
    put("dydx.cpp", f"void dydx_{model}(realtype *dydx, {SIG_X}){{\n" + "\n".join(f"    dydx[{iy + s * ny}] = 1;" for iy, s in enumerate(obs)) + "\n}\n")
    put("sigmay.cpp", f"void sigmay_{model}(realtype *sigmay, const realtype t, const realtype *p, const realtype *k, const realtype *y){{\n" +
        "\n".join(f"    sigmay[{iy}] = {net['sigma']!r};" for iy in range(ny)) + "\n}\n")

    def switch(fn, expr, index):
        body = f"void {fn}_{model}(realtype *{fn}, {SIG_Y}){{\n    switch(iy) {{\n"
        for iy in range(ny):
            body += f"        case {iy}:\n            {fn}[{index(iy)}] = {expr(iy)};\n            break;\n"
        return body + "    }\n}\n"
    put("Jy.cpp", switch("Jy", lambda i: f"0.5*std::log(2*amici::pi*std::pow(sigmay[{i}], 2)) + 0.5*std::pow(-my[{i}] + y[{i}], 2)/std::pow(sigmay[{i}], 2)", lambda i: 0))
    put("dJydsigma.cpp", switch("dJydsigma", lambda i: f"1.0/sigmay[{i}] - 1.0*std::pow(-my[{i}] + y[{i}], 2)/std::pow(sigmay[{i}], 3)", lambda i: i))
    cp = ", ".join("{" + ", ".join("0" if c <= iy else "1" for c in range(ny + 1)) + "}" for iy in range(ny))
    rv = ", ".join("{0}" for _ in range(ny))
    put("dJydy.cpp",
        f"static constexpr std::array<std::array<sunindextype, {ny + 1}>, {ny}> dJydy_colptrs_{model}_ = {{{{ {cp} }}}};\n"
        f"void dJydy_colptrs_{model}(SUNMatrixWrapper &dJydy, int index){{\n    dJydy.set_indexptrs(gsl::make_span(dJydy_colptrs_{model}_[index]));\n}}\n"
        f"static constexpr std::array<std::array<sunindextype, 1>, {ny}> dJydy_rowvals_{model}_ = {{{{ {rv} }}}};\n"
        f"void dJydy_rowvals_{model}(SUNMatrixWrapper &dJydy, int index){{\n    dJydy.set_indexvals(gsl::make_span(dJydy_rowvals_{model}_[index]));\n}}\n" +
        switch("dJydy", lambda i: f"(-1.0*my[{i}] + 1.0*y[{i}])/std::pow(sigmay[{i}], 2)", lambda i: 0))
    # initial state (constants), solver <-> rdata state maps, no splines
    put("x0.cpp", f"void x0_{model}(realtype *x0, const realtype t, const realtype *p, const realtype *k){{\n" +
        "\n".join(f"    x0[{i}] = {float(v)!r};" for i, v in enumerate(net["x0"])) + "\n}\n")
    put("x0_fixedParameters.cpp", f"void x0_fixedParameters_{model}(realtype *x0_fixedParameters, const realtype t, const realtype *p, const realtype *k, gsl::span<const int> reinitialization_state_idxs){{\n}}\n")
    put("sx0_fixedParameters.cpp", f"void sx0_fixedParameters_{model}(realtype *sx0_fixedParameters, const realtype t, const realtype *x0, const realtype *p, const realtype *k, const int ip, gsl::span<const int> reinitialization_state_idxs){{\n}}\n")
    put("x_rdata.cpp", f"void x_rdata_{model}(realtype *x_rdata, const realtype *x, const realtype *tcl, const realtype *p, const realtype *k){{\n    for (int i = 0; i < {nx}; ++i) x_rdata[i] = x[i];\n}}\n")
    put("x_solver.cpp", f"void x_solver_{model}(realtype *x_solver, const realtype *x_rdata){{\n    for (int i = 0; i < {nx}; ++i) x_solver[i] = x_rdata[i];\n}}\n")
    put("create_splines.cpp", '#include "amici/splinefunctions.h"\n#include <vector>\n' +
        PROLOGUE + f"namespace amici {{\nnamespace {ns} {{\nstd::vector<HermiteSpline> create_splines_{model}(const realtype *p, const realtype *k){{\n    return {{}};\n}}\n}}\n}}\n", wrap=False)

    # names
    def names(arr, prefix, n, cnt=None):
        cnt = n if cnt is None else cnt
        return f"extern const std::array<std::string_view const, {cnt}> {arr} = {{\n" + "".join(f'    "{prefix}{i}",\n' for i in range(n)) + "};\n"
    body = "#include <amici/defines.h>\n#include <array>\n#include <string_view>\nnamespace amici::" + ns + " {\n"
    for arr, prefix, n in (("free_parameter_names", "rate_", nr), ("fixed_parameter_names", "k_", 1), ("state_names", "x_", nx),
                           ("state_names_solver", "x_", nx), ("observable_names", "y_", ny), ("free_parameter_ids", "rate_", nr),
                           ("fixed_parameter_ids", "k_", 1), ("state_ids", "x_", nx), ("state_ids_solver", "x_", nx),
                           ("observable_ids", "obs_", ny)):
        body += names(arr, prefix, n)
    body += f"std::array<const ObservableScaling, {ny}> observable_scalings = {{\n" + "    ObservableScaling::lin,\n" * ny + "};\n"
    body += "extern const std::array<std::string_view const, 0> expression_names = {};\nextern const std::array<std::string_view const, 0> expression_ids = {};\n"
    body += f"std::array<int, {nx}> state_idxs_solver = {{\n    " + ", ".join(str(i) for i in range(nx)) + "\n};\n}\n"
    put(f"{model}.cpp", body, wrap=False)
    put("wrapfunctions.h", '#ifndef _amici_wrapfunctions_h\n#define _amici_wrapfunctions_h\n#include <memory>\n#include "amici/model.h"\n'
        "namespace amici::generic_model { std::unique_ptr<amici::Model> get_model(); }\n#endif\n", wrap=False)
    put("wrapfunctions.cpp", f'#include "wrapfunctions.h"\n#include "{model}.h"\n#include "amici/model.h"\nnamespace amici {{\nnamespace generic_model {{\n'
        f"std::unique_ptr<amici::Model> get_model() {{\n    return std::unique_ptr<amici::Model>(new amici::{ns}::Model_{model}());\n}}\n}}\n}}\n", wrap=False)

    # class header: the shipped steadystate header with names / dimensions / defaults substituted
    tmpl = open(os.path.join(ref_root, "models", "model_steadystate_py", "model_steadystate_py.h")).read()
    hdr = tmpl.replace("model_steadystate_py", model)
    sizes = {"free_parameter_names": nr, "fixed_parameter_names": 1, "state_names": nx, "state_names_solver": nx, "observable_names": ny,
             "free_parameter_ids": nr, "fixed_parameter_ids": 1, "state_ids": nx, "state_ids_solver": nx, "observable_ids": ny,
             "observable_scalings": ny, "state_idxs_solver": nx}
    for arr, n in sizes.items():
        hdr, cnt = re.subn(r"(std::array<[^;]*?, )\d+(> " + arr + ";)", r"\g<1>%d\g<2>" % n, hdr)
        assert cnt == 1, arr
    dims = {"nx_rdata": nx, "nxtrue_rdata": nx, "nx_solver": nx, "nxtrue_solver": nx, "np": nr, "nk": 1, "ny": ny, "nytrue": ny,
            "ubw": nx, "lbw": nx, "ndxdotdp_explicit": len(prow), "ndxdotdx_explicit": len(rowval)}
    for key, v in dims.items():
        hdr, cnt = re.subn(r"(\." + key + r" = )\d+,", r"\g<1>%d," % v, hdr)
        assert cnt == 1, key
    hdr, cnt = re.subn(r"\.ndJydy = std::vector<int>\{[^}]*\}", ".ndJydy = std::vector<int>{" + ", ".join("1" for _ in range(ny)) + "}", hdr)
    assert cnt == 1
    hdr, cnt = re.subn(r"std::vector<realtype>\{[^}]*\}, // fixedParameters", "std::vector<realtype>{1.0}, // fixedParameters", hdr)
    assert cnt == 1
    hdr, cnt = re.subn(r"std::vector<realtype>\{[^}]*\}(\s*// free parameters)",
                       "std::vector<realtype>{" + ", ".join(repr(float(v)) for v in net["p_nominal"]) + r"}\g<1>", hdr)
    assert cnt == 1
    hdr, cnt = re.subn(r"std::vector<realtype>\{[^}]*\},(\s*// idlist)", "std::vector<realtype>(%d, 1.0),\\g<1>" % nx, hdr)
    assert cnt == 1
    put(f"{model}.h", hdr, wrap=False)
    return d


if __name__ == "__main__":
    print(write_model(sys.argv[1], sys.argv[2], sys.argv[3]))
