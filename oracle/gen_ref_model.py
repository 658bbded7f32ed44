#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE — writes a ``ModelSpec`` (amici_b200/codegen/emitter.py) as model sources in the format of
the reference's own code generator, so that the UNMODIFIED libamici integrates it with CVODES (+ KLU / dense).

  python oracle/gen_ref_model.py <spec name> <out_dir> <reference root>

Why: BASELINE config 1 is Robertson's problem through the reference's **CVODES** path, but the shipped
``model_robertson_py`` is the index-1 DAE and runs through IDAS (SURVEY.md finding 1, §7 hard part 7).  With this
script ``robertson_ode`` — the same ``ModelSpec`` the CUDA emitter consumes — becomes ``libref_robertson_ode.so`` and the
golden vectors of config 1 come from the integrator the engine must be equivalent to.

Writes <out_dir>/model_<name>_py/{*.cpp, *.h}: one file per model function with the plain-pointer signatures of the
reference's generated code (reference python/sdist/amici/exporters/sundials/cxx_functions.py:101-418; a shipped example
of the layout: models/model_steadystate_py/).  The model class header — boilerplate wiring those functions into
amici::Model_ODE — is derived at build time from the shipped header of a model with the same function set
(models/model_steadystate_py/model_steadystate_py.h, read from the reference tree where it lies) by substituting the model
name, the dimensions and the default parameters; nothing of it is stored in this repository.

Supported: the function set of that header — ODE right-hand side, explicit Jacobian and parameter derivative in CSC form,
observables of the states (no parameter dependence), constant sigmas, initial states that depend on fixed parameters
only, Gaussian objective; no events, splines or expressions.  Anything else raises.
"""
from __future__ import annotations

import os
import re
import sys

import sympy as sp
from sympy.printing.cxx import CXX11CodePrinter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amici_b200.codegen import models as specs  # noqa: E402

PROLOGUE = '#include "amici/symbolic_functions.h"\n#include "amici/defines.h"\n#include "amici/sundials_matrix_wrapper.h"\n' \
           '#include "sundials/sundials_types.h"\n#include <gsl/gsl-lite.hpp>\n#include <algorithm>\n#include <array>\n#include <cmath>\n'
SIG_X = "const realtype t, const realtype *x, const realtype *p, const realtype *k, const realtype *h, const realtype *w"
SIG_Y = "const int iy, const realtype *p, const realtype *k, const realtype *y, const realtype *sigmay, const realtype *my"


class _Cxx(CXX11CodePrinter):
    def _print_Integer(self, expr):
        return f"{float(int(expr))!r}"

    def _print_Rational(self, expr):
        return f"({float(expr.p)!r}/{float(expr.q)!r})"


_pr = _Cxx()


def _index_fn(fn, model, kind, vals):
    name = f"{fn}_{kind}_{model}"
    setter = "set_indexptrs" if kind == "colptrs" else "set_indexvals"
    return (f"static constexpr std::array<sunindextype, {len(vals)}> {name}_ = {{\n    {', '.join(str(v) for v in vals)}\n}};\n"
            f"void {name}(SUNMatrixWrapper &m){{\n    m.{setter}(gsl::make_span({name}_));\n}}\n")


def _csc(mat: sp.Matrix):
    """column pointers, row indices and non-zero expressions of a symbolic matrix (column-major order)"""
    colptr, rowval, vals = [0], [], []
    for c in range(mat.shape[1]):
        for r in range(mat.shape[0]):
            if mat[r, c] != 0:
                rowval.append(r)
                vals.append(mat[r, c])
        colptr.append(len(rowval))
    return colptr, rowval, vals


def write_model(name: str, out_root: str, ref_root: str) -> str:
    spec = specs.ALL[name]()
    if spec.events or spec.splines or spec.event_observables:
        raise NotImplementedError("gen_ref_model: events / splines are outside the function set of the template header")
    model = f"model_{name}_py"
    ns = f"model_{model}"
    d = os.path.join(out_root, model)
    os.makedirs(d, exist_ok=True)
    nx, npar, nk, ny = len(spec.states), len(spec.params), len(spec.fixed), len(spec.observables)
    subs = {s: sp.Symbol(f"x[{i}]") for i, s in enumerate(spec.states)}
    subs.update({s: sp.Symbol(f"p[{i}]") for i, s in enumerate(spec.params)})
    subs.update({s: sp.Symbol(f"k[{i}]") for i, s in enumerate(spec.fixed)})
    pset, kset, xset = set(spec.params), set(spec.fixed), set(spec.states)

    def c(e):
        return _pr.doprint(sp.sympify(e).xreplace(subs))

    def put(fname, body, wrap=True):
        txt = PROLOGUE + f"namespace amici {{\nnamespace {ns} {{\n{body}}} // namespace {ns}\n}} // namespace amici\n" if wrap else body
        path = os.path.join(d, fname)
        if not os.path.exists(path) or open(path).read() != txt:
            with open(path, "w") as f:
                f.write(txt)

    xdot = sp.Matrix([sp.sympify(e) for e in spec.xdot])
    put("xdot.cpp", f"void xdot_{model}(realtype *xdot, {SIG_X}){{\n" +
        "".join(f"    xdot[{i}] = {c(e)};\n" for i, e in enumerate(xdot)) + "}\n")
    J = xdot.jacobian(sp.Matrix(list(spec.states)))
    colptr, rowval, vals = _csc(J)
    put("dxdotdx_explicit.cpp", _index_fn("dxdotdx_explicit", model, "colptrs", colptr) +
        _index_fn("dxdotdx_explicit", model, "rowvals", rowval) +
        f"void dxdotdx_explicit_{model}(realtype *dxdotdx_explicit, {SIG_X}){{\n" +
        "".join(f"    dxdotdx_explicit[{i}] = {c(e)};\n" for i, e in enumerate(vals)) + "}\n")
    P = xdot.jacobian(sp.Matrix(list(spec.params)))
    pcol, prow, pvals = _csc(P)
    put("dxdotdp_explicit.cpp", _index_fn("dxdotdp_explicit", model, "colptrs", pcol) +
        _index_fn("dxdotdp_explicit", model, "rowvals", prow) +
        f"void dxdotdp_explicit_{model}(realtype *dxdotdp_explicit, {SIG_X}){{\n" +
        "".join(f"    dxdotdp_explicit[{i}] = {c(e)};\n" for i, e in enumerate(pvals)) + "}\n")

    # observables: functions of the states only, constant sigma (the template header leaves fdydp / fdsigmaydp empty)
    ys = [sp.sympify(o.formula) for o in spec.observables]
    for o, y in zip(spec.observables, ys):
        if (y.free_symbols & pset) or sp.sympify(o.sigma).free_symbols:
            raise NotImplementedError("gen_ref_model: observables / sigmas that depend on parameters")
    put("y.cpp", f"void y_{model}(realtype *y, {SIG_X}){{\n" + "".join(f"    y[{i}] = {c(e)};\n" for i, e in enumerate(ys)) + "}\n")
    dydx = sp.Matrix(ys).jacobian(sp.Matrix(list(spec.states)))
    put("dydx.cpp", f"void dydx_{model}(realtype *dydx, {SIG_X}){{\n" +
        "".join(f"    dydx[{iy + ix * ny}] = {c(dydx[iy, ix])};\n" for ix in range(nx) for iy in range(ny) if dydx[iy, ix] != 0) + "}\n")
    put("sigmay.cpp", f"void sigmay_{model}(realtype *sigmay, const realtype t, const realtype *p, const realtype *k, const realtype *y){{\n" +
        "".join(f"    sigmay[{iy}] = {float(o.sigma)!r};\n" for iy, o in enumerate(spec.observables)) + "}\n")

    def switch(fn, expr, index):
        body = f"void {fn}_{model}(realtype *{fn}, {SIG_Y}){{\n    switch(iy) {{\n"
        for iy in range(ny):
            body += f"        case {iy}:\n            {fn}[{index(iy)}] = {expr(iy)};\n            break;\n"
        return body + "    }\n}\n"
    put("Jy.cpp", switch("Jy", lambda i: f"0.5*std::log(2*amici::pi*std::pow(sigmay[{i}], 2)) + 0.5*std::pow(-my[{i}] + y[{i}], 2)/std::pow(sigmay[{i}], 2)", lambda i: 0))
    put("dJydsigma.cpp", switch("dJydsigma", lambda i: f"1.0/sigmay[{i}] - 1.0*std::pow(-my[{i}] + y[{i}], 2)/std::pow(sigmay[{i}], 3)", lambda i: i))
    cp = ", ".join("{" + ", ".join("0" if cc <= iy else "1" for cc in range(ny + 1)) + "}" for iy in range(ny))
    rv = ", ".join("{0}" for _ in range(ny))
    put("dJydy.cpp",
        f"static constexpr std::array<std::array<sunindextype, {ny + 1}>, {ny}> dJydy_colptrs_{model}_ = {{{{ {cp} }}}};\n"
        f"void dJydy_colptrs_{model}(SUNMatrixWrapper &dJydy, int index){{\n    dJydy.set_indexptrs(gsl::make_span(dJydy_colptrs_{model}_[index]));\n}}\n"
        f"static constexpr std::array<std::array<sunindextype, 1>, {ny}> dJydy_rowvals_{model}_ = {{{{ {rv} }}}};\n"
        f"void dJydy_rowvals_{model}(SUNMatrixWrapper &dJydy, int index){{\n    dJydy.set_indexvals(gsl::make_span(dJydy_rowvals_{model}_[index]));\n}}\n" +
        switch("dJydy", lambda i: f"(-1.0*my[{i}] + 1.0*y[{i}])/std::pow(sigmay[{i}], 2)", lambda i: 0))

    # initial state: may depend on fixed parameters (x0_fixedParameters), not on parameters (fsx0 is empty in the template)
    x0 = [sp.sympify(e) for e in spec.x0]
    if any(e.free_symbols & (pset | xset) for e in x0):
        raise NotImplementedError("gen_ref_model: initial states that depend on parameters")
    put("x0.cpp", f"void x0_{model}(realtype *x0, const realtype t, const realtype *p, const realtype *k){{\n" +
        "".join(f"    x0[{i}] = {c(e)};\n" for i, e in enumerate(x0) if e != 0) + "}\n")
    fixed_idx = [i for i, e in enumerate(x0) if e.free_symbols & kset]
    find = "std::find(reinitialization_state_idxs.cbegin(), reinitialization_state_idxs.cend(), {i}) != reinitialization_state_idxs.cend()"
    put("x0_fixedParameters.cpp", f"void x0_fixedParameters_{model}(realtype *x0_fixedParameters, const realtype t, const realtype *p, const realtype *k, gsl::span<const int> reinitialization_state_idxs){{\n" +
        "".join(f"    if({find.format(i=i)})\n        x0_fixedParameters[{i}] = {c(x0[i])};\n" for i in fixed_idx) + "}\n")
    body = f"void sx0_fixedParameters_{model}(realtype *sx0_fixedParameters, const realtype t, const realtype *x0, const realtype *p, const realtype *k, const int ip, gsl::span<const int> reinitialization_state_idxs){{\n"
    if fixed_idx:
        body += (f"    static const std::array<int, {len(fixed_idx)}> _x0_fixedParameters_idxs = {{\n        {', '.join(map(str, fixed_idx))}\n    }};\n"
                 "    for(auto idx: reinitialization_state_idxs) {\n"
                 "        if(std::find(_x0_fixedParameters_idxs.cbegin(), _x0_fixedParameters_idxs.cend(), idx) != _x0_fixedParameters_idxs.cend())\n"
                 "            sx0_fixedParameters[idx] = 0.0;\n    }\n")
    put("sx0_fixedParameters.cpp", body + "}\n")
    put("x_rdata.cpp", f"void x_rdata_{model}(realtype *x_rdata, const realtype *x, const realtype *tcl, const realtype *p, const realtype *k){{\n    for (int i = 0; i < {nx}; ++i) x_rdata[i] = x[i];\n}}\n")
    put("x_solver.cpp", f"void x_solver_{model}(realtype *x_solver, const realtype *x_rdata){{\n    for (int i = 0; i < {nx}; ++i) x_solver[i] = x_rdata[i];\n}}\n")
    put("create_splines.cpp", '#include "amici/splinefunctions.h"\n#include <vector>\n' +
        PROLOGUE + f"namespace amici {{\nnamespace {ns} {{\nstd::vector<HermiteSpline> create_splines_{model}(const realtype *p, const realtype *k){{\n    return {{}};\n}}\n}}\n}}\n", wrap=False)

    def names(arr, items):
        return f"extern const std::array<std::string_view const, {len(items)}> {arr} = {{\n" + "".join(f'    "{s}",\n' for s in items) + "};\n"
    pn, kn, xn = [str(s) for s in spec.params], [str(s) for s in spec.fixed], [str(s) for s in spec.states]
    on = [o.name for o in spec.observables]
    body = "#include <amici/defines.h>\n#include <array>\n#include <string_view>\nnamespace amici::" + ns + " {\n"
    for arr, items in (("free_parameter_names", pn), ("fixed_parameter_names", kn), ("state_names", xn), ("state_names_solver", xn),
                       ("observable_names", on), ("free_parameter_ids", pn), ("fixed_parameter_ids", kn), ("state_ids", xn),
                       ("state_ids_solver", xn), ("observable_ids", on)):
        body += names(arr, items)
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
    sizes = {"free_parameter_names": npar, "fixed_parameter_names": nk, "state_names": nx, "state_names_solver": nx,
             "observable_names": ny, "free_parameter_ids": npar, "fixed_parameter_ids": nk, "state_ids": nx,
             "state_ids_solver": nx, "observable_ids": ny, "observable_scalings": ny, "state_idxs_solver": nx}
    for arr, n in sizes.items():
        hdr, cnt = re.subn(r"(std::array<[^;]*?, )\d+(> " + arr + ";)", r"\g<1>%d\g<2>" % n, hdr)
        assert cnt == 1, arr
    dims = {"nx_rdata": nx, "nxtrue_rdata": nx, "nx_solver": nx, "nxtrue_solver": nx, "nx_solver_reinit": len(fixed_idx), "np": npar,
            "nk": nk, "ny": ny, "nytrue": ny, "ubw": nx, "lbw": nx, "ndxdotdp_explicit": len(prow), "ndxdotdx_explicit": len(rowval)}
    for key, v in dims.items():
        hdr, cnt = re.subn(r"(\." + key + r" = )\d+,", r"\g<1>%d," % v, hdr)
        assert cnt == 1, key
    hdr, cnt = re.subn(r"\.ndJydy = std::vector<int>\{[^}]*\}", ".ndJydy = std::vector<int>{" + ", ".join("1" for _ in range(ny)) + "}", hdr)
    assert cnt == 1
    hdr, cnt = re.subn(r"std::vector<realtype>\{[^}]*\}, // fixedParameters",
                       "std::vector<realtype>{" + ", ".join(repr(float(v)) for v in spec.k_nominal) + "}, // fixedParameters", hdr)
    assert cnt == 1
    hdr, cnt = re.subn(r"std::vector<realtype>\{[^}]*\}(\s*// free parameters)",
                       "std::vector<realtype>{" + ", ".join(repr(float(v)) for v in spec.p_nominal) + r"}\g<1>", hdr)
    assert cnt == 1
    hdr, cnt = re.subn(r"std::vector<realtype>\{[^}]*\},(\s*// idlist)", "std::vector<realtype>(%d, 1.0),\\g<1>" % nx, hdr)
    assert cnt == 1
    put(f"{model}.h", hdr, wrap=False)
    return d


if __name__ == "__main__":
    print(write_model(sys.argv[1], sys.argv[2], sys.argv[3]))
