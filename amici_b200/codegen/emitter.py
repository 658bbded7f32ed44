"""CUDA emitter: symbolic model description -> inlined FP64 ``__host__ __device__`` functions.

This is the B200 counterpart of the reference's C++ code generator
(reference python/sdist/amici/exporters/sundials/de_export.py:267-508 writes one
``<fn>_<model>.cpp`` per model function with the signatures of
exporters/sundials/cxx_functions.py:101-418).  The reference back end cannot be
imported in this image (libsbml/toposort are missing), so the emitter starts from
a small SymPy model description (``ModelSpec``) instead of a ``DEModel``; the set
of functions it writes is the one the batched BDF engine consumes.

Differences from the reference's generated code that are deliberate:

* One header per model holding a traits struct with ``static`` inline functions;
  no sparse-matrix objects at run time (the CSC pattern is burnt into straight
  line code: ``jac`` writes the non-zeros, ``jac_mul_add``/``jac_to_A`` are
  unrolled over the static pattern).
* ``sxdot`` pieces are *column functions*: ``dxdotdp(ip, ...)`` returns the total
  derivative column for one parameter.  Instead of the reference's
  ``switch(ip)`` (dydp.cpp, sx0.cpp ...) they are select chains
  ``(ip == 3) ? e3 : (ip == 7) ? e7 : 0`` so that the lanes of a warp, each
  holding a different parameter column, do not diverge.
* ``w`` expressions are inlined; only spline outputs stay opaque (``spl[i]``,
  ``sspl[i]`` = d spl_i / d p_ip for the calling lane's parameter).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Sequence

import sympy as sp
from sympy.printing.c import C99CodePrinter


@dataclasses.dataclass
class Spline:
    """Cubic Hermite spline in time (reference src/splinefunctions.cpp, HermiteSpline).

    Supported configuration = what the reference's shipped models use
    (models/model_jakstat_adjoint_py/create_splines.cpp): non-equidistant nodes,
    node derivatives by finite differences, zero-derivative boundary conditions,
    constant extrapolation, optional logarithmic parametrisation.
    """
    name: str
    nodes: Sequence[float]
    values: Sequence[sp.Expr]
    logarithmic: bool = True
    # the rest of the reference's HermiteSpline configuration (csrc/hermite_spline.cuh); defaults = the shipped models'
    derivatives: Sequence[sp.Expr] | None = None      # node slopes; None: finite differences
    bc: tuple = ("zeroDerivative", "zeroDerivative")   # given | zeroDerivative | natural | periodic
    extrapolate: tuple = ("constant", "constant")      # noExtrapolation | constant | linear | polynomial | periodic
    equidistant: bool = False                           # nodes = linspace(nodes[0], nodes[-1], len(values))

    def is_default(self) -> bool:
        return (self.derivatives is None and tuple(self.bc) == ("zeroDerivative", "zeroDerivative")
                and tuple(self.extrapolate) == ("constant", "constant") and not self.equidistant)

    def node_list(self) -> list:
        if not self.equidistant:
            return [float(v) for v in self.nodes]
        n = len(self.values)
        step = (float(self.nodes[-1]) - float(self.nodes[0])) / (n - 1)    # reference src/splinefunctions.cpp:40-51
        return [float(self.nodes[0]) + i * step for i in range(n - 1)] + [float(self.nodes[-1])]


@dataclasses.dataclass
class Observable:
    name: str
    formula: sp.Expr
    sigma: sp.Expr = sp.Float(1.0)


@dataclasses.dataclass
class Event:
    """AMICI event (reference include/amici/event.h): fires when ``root`` crosses zero from below.

    ``root`` is the trigger function g(t, x, p, k).  Events whose root depends on the state are
    tracked by the integrator's root finder and must come first (``ne_solver`` of the reference);
    events whose root depends only on time and parameters (the Heaviside/piecewise switches the
    reference importer creates in +/- pairs, cf. models/model_events_py/root.cpp) are handled as
    explicit stop times.  ``assignments`` maps state symbols to their new value (evaluated on the
    pre-event state).
    """
    root: sp.Expr
    assignments: dict | None = None
    initial_value: bool = True


@dataclasses.dataclass
class EventObservable:
    """Event-resolved observable z recorded when event ``event`` fires (MeasurementChannel with event_id)."""
    name: str
    event: int
    formula: sp.Expr
    sigma: sp.Expr = sp.Float(1.0)


@dataclasses.dataclass
class ModelSpec:
    name: str
    states: Sequence[sp.Symbol]
    params: Sequence[sp.Symbol]
    fixed: Sequence[sp.Symbol]
    xdot: Sequence[sp.Expr]
    x0: Sequence[sp.Expr]
    observables: Sequence[Observable]
    p_nominal: Sequence[float]
    k_nominal: Sequence[float]
    splines: Sequence[Spline] = ()
    events: Sequence[Event] = ()
    event_observables: Sequence[EventObservable] = ()
    heavisides: dict | None = None         # {symbol used in xdot/y: event index}: value of h[event index]
    algebraic: Sequence[bool] = ()         # DAE models: True for algebraic states; their `xdot` entry is the algebraic equation
                                           # 0 = g(t, x, p, k) (mass matrix diag(not algebraic), reference src/model_dae.cpp)
    doc: str = ""


t_sym = sp.Symbol("t", real=True)


class _Printer(C99CodePrinter):
    """C printer: small integer powers -> products, everything double."""

    def _print_Pow(self, expr):
        b, e = expr.as_base_exp()
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self._print(b)
            if not (b.is_Symbol or b.is_Number):
                s = f"({s})"
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self._print(b)
            if not (b.is_Symbol or b.is_Number):
                s = f"({s})"
            return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
        return super()._print_Pow(expr)

    def _print_Rational(self, expr):
        return f"({float(expr.p)!r}/{float(expr.q)!r})"

    def _print_Integer(self, expr):
        return f"{float(int(expr))!r}"

    def _print_Heaviside(self, expr):
        return f"((({self._print(expr.args[0])}) > 0.0) ? 1.0 : 0.0)"


_printer = _Printer()


def _c(e) -> str:
    return _printer.doprint(sp.sympify(e))


class _FnWriter:
    """Writes one function body: CSE over all outputs, then assignments.

    ``hoist=True`` (used for the functions of the integrator's inner loop: xdot, jac, dxdotdp) moves every maximal
    sub-expression that depends on parameters only out of the function: it becomes a *derived constant*, evaluated
    once per condition by the generated ``derived()`` and stored behind the fixed parameters (``k[NK + i]``).
    The reference's generated code re-evaluates such terms (e.g. the compartment-volume quotients of JAK-STAT,
    an FP64 division each) in every right-hand-side call.
    """

    def __init__(self, subs: dict, const_syms=(), nk: int = 0):
        self.subs = subs  # symbol -> C lvalue text
        self.const_syms = set(const_syms)   # parameters + fixed parameters
        self.nk = nk
        self.derived: list[sp.Expr] = []    # derived constant i lives in k[nk + i]

    def _derived_symbol(self, e):
        for i, d in enumerate(self.derived):
            if d == e:
                return sp.Symbol(f"k[{self.nk + i}]")
        self.derived.append(e)
        return sp.Symbol(f"k[{self.nk + len(self.derived) - 1}]")

    def _is_const(self, e):
        return bool(e.free_symbols) and e.free_symbols <= self.const_syms

    def _worth(self, e):
        # a bare symbol or its negation costs nothing at the point of use
        if e.is_Atom:
            return False
        if e.is_Mul and len(e.args) == 2 and e.args[0] == -1 and e.args[1].is_Atom:
            return False
        return True

    def _hoist(self, e):
        e = sp.sympify(e)
        if e.is_Atom:
            return e
        if self._is_const(e):
            return self._derived_symbol(e) if self._worth(e) else e
        if e.is_Mul or e.is_Add:
            cargs = [a for a in e.args if self._is_const(a)]
            nums = [a for a in e.args if a.is_Number]
            rest = [self._hoist(a) for a in e.args if not self._is_const(a) and not a.is_Number]
            if cargs:
                c = e.func(*(cargs + (nums if e.is_Mul else [])))
                head = self._derived_symbol(c) if self._worth(c) else c
                tail = [] if e.is_Mul else nums
                return e.func(head, *tail, *rest)
            return e.func(*nums, *rest)
        return e.func(*[self._hoist(a) for a in e.args])

    def body(self, assigns: list[tuple[str, sp.Expr]], indent="        ", hoist: bool = False) -> str:
        if not assigns:
            return ""
        exprs = [sp.sympify(e) for _, e in assigns]
        if hoist:
            exprs = [self._hoist(sp.expand_mul(e) if False else e) for e in exprs]
        repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("c_"), optimizations="basic")
        lines = []
        for s, e in repl:
            lines.append(f"{indent}const double {s} = {self._pr(e)};")
        for (lhs, _), e in zip(assigns, red):
            lines.append(f"{indent}{lhs} = {self._pr(e)};")
        return "\n".join(lines) + "\n"

    def _pr(self, e) -> str:
        e = sp.sympify(e)
        return _c(e.xreplace(self.subs))


def _select_chain(ip_var: str, cases: list[tuple[int, str]]) -> str:
    """(ip==a)? ea : (ip==b)? eb : 0.0"""
    out = "0.0"
    for ip, txt in reversed(cases):
        if isinstance(ip, (list, tuple)):   # several parameters share one value
            ips = sorted(ip)
            if len(ips) > 2 and ips == list(range(ips[0], ips[-1] + 1)):
                cond = f"({ip_var} >= {ips[0]} && {ip_var} <= {ips[-1]})"
            else:
                cond = "(" + " || ".join(f"{ip_var} == {i}" for i in ips) + ")"
        else:
            cond = f"({ip_var} == {ip})"
        out = f"{cond} ? ({txt}) : {out}"
    return out


def emit_model(spec: ModelSpec) -> str:
    nx, npar, nk, ny = len(spec.states), len(spec.params), len(spec.fixed), len(spec.observables)
    nspl, ne = len(spec.splines), len(spec.events)
    spl_syms = [sp.Symbol(f"spl_{s.name}", real=True) for s in spec.splines]
    sspl_syms = [sp.Symbol(f"sspl_{s.name}", real=True) for s in spec.splines]
    spl_map = {sp.Symbol(s.name, real=True): spl_syms[i] for i, s in enumerate(spec.splines)}
    spl_map.update({sp.Symbol(s.name): spl_syms[i] for i, s in enumerate(spec.splines)})
    h_map = dict(spec.heavisides or {})

    subs = {}
    for i, s in enumerate(spec.states):
        subs[s] = sp.Symbol(f"x[{i}]")
    for i, s in enumerate(spec.params):
        subs[s] = sp.Symbol(f"p[{i}]")
    for i, s in enumerate(spec.fixed):
        subs[s] = sp.Symbol(f"k[{i}]")
    for i, s in enumerate(spl_syms):
        subs[s] = sp.Symbol(f"spl[{i}]")
    for i, s in enumerate(sspl_syms):
        subs[s] = sp.Symbol(f"sspl[{i}]")
    for s, ie in h_map.items():
        subs[s] = sp.Symbol(f"h[{ie}]")
    y_syms = [sp.Symbol(f"y_{o.name}", real=True) for o in spec.observables]
    sig_syms = [sp.Symbol(f"sigma_{o.name}", positive=True) for o in spec.observables]
    my_syms = [sp.Symbol(f"my_{o.name}", real=True) for o in spec.observables]
    for i in range(ny):
        subs[y_syms[i]] = sp.Symbol(f"y[{i}]")
        subs[sig_syms[i]] = sp.Symbol(f"sigmay[{i}]")
        subs[my_syms[i]] = sp.Symbol(f"my[{i}]")
    sx_syms = [sp.Symbol(f"sx_{i}", real=True) for i in range(nx)]
    for i in range(nx):
        subs[sx_syms[i]] = sp.Symbol(f"sx[{i}]")
    W = _FnWriter(subs, const_syms=list(spec.params) + list(spec.fixed), nk=nk)

    xdot = [sp.sympify(e).xreplace(spl_map) for e in spec.xdot]
    x0 = [sp.sympify(e) for e in spec.x0]
    yexpr = [sp.sympify(o.formula).xreplace(spl_map) for o in spec.observables]
    sigexpr = [sp.sympify(o.sigma) for o in spec.observables]

    # ---- Jacobian (static CSC pattern, column major like the reference's dxdotdx) ----
    Jfull = [[sp.diff(xdot[r], spec.states[c]) for c in range(nx)] for r in range(nx)]
    colptr, rowval, jvals = [0], [], []
    for c in range(nx):
        for r in range(nx):
            if Jfull[r][c] != 0:
                rowval.append(r)
                jvals.append(Jfull[r][c])
        colptr.append(len(rowval))
    nnz = len(rowval)

    # ---- total parameter derivative columns ----
    def total_dp(expr, ip):
        e = sp.diff(expr, spec.params[ip])
        for i, s in enumerate(spl_syms):
            # d spl_i / d p_ip is supplied at run time (sspl[i]) for the lane's own parameter
            if spec.splines[i] and any(sp.diff(sp.sympify(v), spec.params[ip]) != 0 for v in list(spec.splines[i].values) + list(spec.splines[i].derivatives or [])):
                e = e + sp.diff(expr, s) * sspl_syms[i]
        return sp.simplify(e) if e != 0 else e

    out = []
    A = out.append
    guard = f"AMICI_B200_MODEL_{spec.name.upper()}_CUH"
    origin = spec.doc if spec.doc.startswith("from ") else f"from amici_b200/codegen/models.py::{spec.name}"
    A(f"// GENERATED by amici_b200/codegen/emitter.py {origin} — do not edit.")
    A(f"// Device functions of model '{spec.name}' for the batched BDF engine (see emitter.py for the")
    A("// mapping onto the reference's generated <fn>_<model>.cpp files).")
    A(f"#ifndef {guard}\n#define {guard}\n#include \"../hermite_spline.cuh\"\n#include \"../model_common.cuh\"\nnamespace amici_b200 {{ namespace models {{\n")
    A(f"struct {spec.name} {{")
    A(f"    static constexpr int NX = {nx}, NP = {npar}, NK = {nk}, NY = {ny}, NSPL = {nspl}, NE = {ne}, NH = {ne};")
    A(f"    static constexpr int NNZ_J = {nnz};")
    nz = len(spec.event_observables)
    A(f"    static constexpr int NZ = {nz};")
    nn_max = max([len(s.nodes) for s in spec.splines], default=1)
    A(f"    static constexpr int SPL_NN = {nn_max};  // nodes per spline")
    A(f"    static constexpr bool SPL_LOG = {'true' if all(s.logarithmic for s in spec.splines) else 'false'};")
    spl_general = any(not s.is_default() for s in spec.splines) or len({bool(s.logarithmic) for s in spec.splines}) > 1
    A(f"    static constexpr bool SPL_GENERAL = {'true' if spl_general else 'false'};   // any spline outside the default HermiteSpline configuration")
    A(f"    static const char* name() {{ return \"{spec.name}\"; }}")
    A("    static void nominal(double* p, double* k) {")
    for i, v in enumerate(spec.p_nominal):
        A(f"        p[{i}] = {float(v)!r};")
    for i, v in enumerate(spec.k_nominal):
        A(f"        k[{i}] = {float(v)!r};")
    A("    }")

    sig_common = "const double t, const double* x, const double* p, const double* k, const double* h, const double* spl"

    # x0 / sx0
    # states whose initial value depends on fixed parameters (the index list of the reference's generated x0_fixedParameters.cpp);
    # the reference disables re-initialisation at import when such a state also depends on parameters
    kset, pset = set(spec.fixed), set(spec.params)
    fixed_mask = sum(1 << i for i in range(nx) if sp.sympify(x0[i]).free_symbols & kset)
    allowed = not any(sp.sympify(x0[i]).free_symbols & kset and sp.sympify(x0[i]).free_symbols & pset for i in range(nx))
    A(f"    static constexpr unsigned long long X0_FIXED_MASK = {fixed_mask:#x}ull;")
    A(f"    static constexpr bool X0_FIXED_REINIT_ALLOWED = {'true' if allowed else 'false'};")
    # Parameters whose state-sensitivity column is identically zero: no entry of xdot, of the initial state or of a spline
    # depends on them (observable offsets / scalings / noise parameters) and the model has no events.  For such a column
    # sx0 = 0 and d/dt sx = J*sx + 0, so CVODES carries exact zeros that never win a max-norm: the engine can leave the
    # column out of the integration (Layout<.., COMPACT>) without changing a single step-control decision; their output
    # sensitivities (sy = dy/dp, ssigmay, sres, FIM, sllh) are formed from sx = 0 at the output points.
    static_mask = 0
    if ne == 0 and npar <= 64:
        for ip in range(npar):
            par = spec.params[ip]
            dyn = any(sp.diff(e, par) != 0 for e in xdot) or any(sp.diff(e, par) != 0 for e in x0) or \
                any(sp.diff(sp.sympify(v), par) != 0 for spl in spec.splines for v in list(spl.values) + list(spl.derivatives or []))
            if not dyn:
                static_mask |= 1 << ip
    A(f"    static constexpr unsigned long long SX_STATIC_MASK = {static_mask:#x}ull;   // parameters with sx == 0 identically")
    A(f"    static constexpr int NP_DYN = {npar - bin(static_mask).count('1')};   // parameters that need an integrated sensitivity column")
    alg_mask = sum(1 << i for i, a in enumerate(spec.algebraic) if a)
    if alg_mask:
        A(f"    static constexpr unsigned long long ALG_MASK = {alg_mask:#x}ull;   // algebraic states (DAE): row r of xdot is 0 = g_r")
    A("    AB_HD static void x0(const double* p, const double* k, double* x0) {")
    A("        (void)p; (void)k;")
    A(W.body([(f"x0[{i}]", x0[i]) for i in range(nx)]).rstrip("\n"))
    A("    }")
    A("    AB_HD static void sx0(const int ip, const double* p, const double* k, double* sx0) {")
    A("        (void)p; (void)k; (void)ip;")
    for i in range(nx):
        cases = [(ip, W._pr(sp.diff(x0[i], spec.params[ip]))) for ip in range(npar) if sp.diff(x0[i], spec.params[ip]) != 0]
        A(f"        sx0[{i}] = {_select_chain('ip', cases)};")
    A("    }")

    # xdot
    A(f"    AB_HD static void xdot({sig_common}, double* xdot) {{")
    A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl;")
    A(W.body([(f"xdot[{i}]", xdot[i]) for i in range(nx)], hoist=True).rstrip("\n"))
    A("    }")
    # jac
    A(f"    // CSC pattern of J = d xdot / d x (cf. reference dxdotdx_explicit_{{colptrs,rowvals}})")
    A(f"    // colptr = {colptr}\n    // rowval = {rowval}")
    A(f"    AB_HD static void jac({sig_common}, double* J) {{")
    A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl;")
    A(W.body([(f"J[{i}]", jvals[i]) for i in range(nnz)], hoist=True).rstrip("\n"))
    A("    }")
    A("    // out += J * s")
    A("    AB_HD static void jac_mul_add(const double* J, const double* s, double* out) {")
    for r in range(nx):
        terms = []
        for c in range(nx):
            for idx in range(colptr[c], colptr[c + 1]):
                if rowval[idx] == r:
                    terms.append(f"J[{idx}] * s[{c}]")
        if terms:
            A(f"        out[{r}] += " + " + ".join(terms) + ";")
    A("    }")
    A("    // A (row-major NX*NX) = I - gamma * J")
    A("    template <class Mat> AB_HD static void jac_to_A(const double* J, const double gamma, Mat& A) {")
    for r in range(nx):
        for c in range(nx):
            idx = None
            for q in range(colptr[c], colptr[c + 1]):
                if rowval[q] == r:
                    idx = q
            diag = "1.0" if r == c else "0.0"
            if idx is None:
                A(f"        A({r}, {c}) = {diag};")
            else:
                A(f"        A({r}, {c}) = {diag} - gamma * J[{idx}];")
    A("    }")
    A("    // dense row-major copy of J (ReturnData::J layout, reference include/amici/rdata.h:707-728)")
    A("    AB_HD static void jac_to_dense(const double* J, double* D) {")
    for r in range(nx):
        for c in range(nx):
            idx = None
            for q in range(colptr[c], colptr[c + 1]):
                if rowval[q] == r:
                    idx = q
            A(f"        D[{r * nx + c}] = " + (f"J[{idx}];" if idx is not None else "0.0;"))
    A("    }")

    # ---- sparse LU of A = I - gamma*J with a static, fill-reducing diagonal pivot order ----
    # (replaces SUNLinSol KLU/dense for small systems: straight-line factor and solve code; the run time
    #  checks KLU's diagonal-preference test |a_kk| >= 1e-3 * max|a_ik| and falls back to dense partial
    #  pivoting when it fails, reference ThirdParty/sundials/src/sunlinsol/klu/sunlinsol_klu.c:240-342)
    pat = [[(r == c) or (Jfull[r][c] != 0) for c in range(nx)] for r in range(nx)]
    jidx = {}
    for c in range(nx):
        for q_ in range(colptr[c], colptr[c + 1]):
            jidx[(rowval[q_], c)] = q_
    work = [row[:] for row in pat]
    remaining = list(range(nx))
    order = []
    while remaining:
        best, bestcost = None, None
        for kk in remaining:
            rc = sum(1 for i in remaining if work[i][kk]) - 1
            cc_ = sum(1 for j in remaining if work[kk][j]) - 1
            cost = rc * cc_
            if bestcost is None or cost < bestcost:
                best, bestcost = kk, cost
        kk = best
        remaining.remove(kk)
        order.append(kk)
        for i in remaining:
            if work[i][kk]:
                for j in remaining:
                    if work[kk][j]:
                        work[i][j] = True
    # symbolic elimination in that order, emitting code
    fact, solve_f, solve_b = [], [], []
    cur = [row[:] for row in pat]
    done = set()
    lu_slots = {}

    def slot(name):
        if name not in lu_slots:
            lu_slots[name] = len(lu_slots)
        return lu_slots[name]

    for r in range(nx):
        for c in range(nx):
            if pat[r][c]:
                diag = "1.0" if r == c else "0.0"
                if (r, c) in jidx:
                    fact.append(f"double a_{r}_{c} = {diag} - gamma * J[{jidx[(r, c)]}];")
                else:
                    fact.append(f"double a_{r}_{c} = {diag};")
    declared = {(r, c) for r in range(nx) for c in range(nx) if pat[r][c]}
    fact.append("bool ok = true;")
    for kk in order:
        rows_k = [i for i in range(nx) if i not in done and i != kk and cur[i][kk]]
        cols_k = [j for j in range(nx) if j not in done and j != kk and cur[kk][j]]
        if rows_k:
            mx = " , ".join(f"fabs(a_{i}_{kk})" for i in rows_k)
            expr = f"fabs(a_{rows_k[0]}_{kk})"
            for i in rows_k[1:]:
                expr = f"ab_max({expr}, fabs(a_{i}_{kk}))"
            fact.append(f"ok = ok && (fabs(a_{kk}_{kk}) >= 1e-3 * ({expr}));")
        else:
            fact.append(f"ok = ok && (a_{kk}_{kk} != 0.0);")
        fact.append(f"const double inv_{kk} = fast_rcp(a_{kk}_{kk});")
        fact.append(f"LU[{slot(f'inv_{kk}')}] = inv_{kk};")
        for i in rows_k:
            fact.append(f"const double l_{i}_{kk} = a_{i}_{kk} * inv_{kk};")
            fact.append(f"LU[{slot(f'l_{i}_{kk}')}] = l_{i}_{kk};")
            solve_f.append(f"b[{i}] -= LU[{slot(f'l_{i}_{kk}')}] * b[{kk}];")
            for j in cols_k:
                if (i, j) in declared:
                    fact.append(f"a_{i}_{j} -= l_{i}_{kk} * a_{kk}_{j};")
                else:
                    fact.append(f"double a_{i}_{j} = -l_{i}_{kk} * a_{kk}_{j};")
                    declared.add((i, j))
                cur[i][j] = True
        for j in cols_k:
            fact.append(f"LU[{slot(f'u_{kk}_{j}')}] = a_{kk}_{j};")
        terms = "".join(f" - LU[{slot(f'u_{kk}_{j}')}] * b[{j}]" for j in cols_k)
        solve_b.append(f"b[{kk}] = (b[{kk}]{terms}) * LU[{slot(f'inv_{kk}')}];")
        done.add(kk)
    nlu = len(lu_slots)
    A(f"    static constexpr int NLU = {nlu};  // stored entries of the sparse LU (pivot order {order})")
    A("    // LU <- factors of I - gamma*J (no row exchanges); returns false if a pivot fails the diagonal-preference test")
    A("    AB_HD static bool lu_factor(const double* J, const double gamma, double* LU) {")
    for ln in fact:
        A("        " + ln)
    A("        return ok;")
    A("    }")
    A("    // b <- (I - gamma*J)^{-1} b")
    A("    AB_HD static void lu_solve(const double* LU, double* b) {")
    for ln in solve_f:
        A("        " + ln)
    for ln in reversed(solve_b):
        A("        " + ln)
    A("    }")

    # dxdotdp column
    A("    // total derivative column d xdot / d p_ip (explicit + through splines), select-chain form")
    A(f"    AB_HD static void dxdotdp(const int ip, {sig_common}, const double* sspl, double* out) {{")
    A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)sspl; (void)ip;")
    vals, rows = [], []
    for r in range(nx):
        groups: list[tuple[sp.Expr, list[int]]] = []   # parameters with identical derivative share one select
        for ip in range(npar):
            e = total_dp(xdot[r], ip)
            if e == 0:
                continue
            for ge, gi in groups:
                if ge == e:
                    gi.append(ip)
                    break
            else:
                groups.append((e, [ip]))
        cases = []
        for e, ips in groups:
            vals.append((f"const double v_{r}_{ips[0]}", e))
            cases.append((ips if len(ips) > 1 else ips[0], f"v_{r}_{ips[0]}"))
        rows.append(cases)
    A(W.body(vals, hoist=True).rstrip("\n"))
    for r in range(nx):
        A(f"        out[{r}] = {_select_chain('ip', rows[r])};")
    A("    }")

    # observables
    A(f"    AB_HD static void y({sig_common}, double* y) {{")
    A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl;")
    A(W.body([(f"y[{i}]", yexpr[i]) for i in range(ny)]).rstrip("\n"))
    A("    }")
    A("    // sy = dydx * sx + dydp(ip)   (reference Model::get_observable_sensitivity, src/model.cpp:1123-1156)")
    A(f"    AB_HD static void sy(const int ip, {sig_common}, const double* sspl, const double* sx, double* sy) {{")
    A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)sspl; (void)ip; (void)sx;")
    vals, rows = [], []
    for iy in range(ny):
        lin = sum((sp.diff(yexpr[iy], spec.states[c]) * sx_syms[c] for c in range(nx)), sp.Integer(0))
        vals.append((f"const double l_{iy}", lin))
        cases = []
        for ip in range(npar):
            e = total_dp(yexpr[iy], ip)
            if e != 0:
                vals.append((f"const double v_{iy}_{ip}", e))
                cases.append((ip, f"v_{iy}_{ip}"))
        rows.append(cases)
    A(W.body(vals).rstrip("\n"))
    for iy in range(ny):
        A(f"        sy[{iy}] = l_{iy} + ({_select_chain('ip', rows[iy])});")
    A("    }")
    ysub = {}
    # sigma may depend on p, k and y (symbols y_<name>)
    A("    AB_HD static void sigmay(const double* p, const double* k, const double* y, double* sigmay) {")
    A("        (void)p; (void)k; (void)y;")
    A(W.body([(f"sigmay[{i}]", sigexpr[i]) for i in range(ny)]).rstrip("\n"))
    A("    }")
    A("    // ssigmay = dsigmaydp(ip) + dsigmaydy * sy   (reference src/model.cpp:1165-1192)")
    A("    AB_HD static void ssigmay(const int ip, const double* p, const double* k, const double* y, const double* sy, double* out) {")
    A("        (void)p; (void)k; (void)y; (void)sy; (void)ip;")
    sy_syms = [sp.Symbol(f"sy[{i}]") for i in range(ny)]
    for iy in range(ny):
        cases = [(ip, W._pr(sp.diff(sigexpr[iy], spec.params[ip]))) for ip in range(npar)
                 if sp.diff(sigexpr[iy], spec.params[ip]) != 0]
        lin = sum((sp.diff(sigexpr[iy], y_syms[j]) * sy_syms[j] for j in range(ny)), sp.Integer(0))
        A(f"        out[{iy}] = ({_select_chain('ip', cases)}) + {W._pr(lin)};")
    A("    }")
    A("    // only the explicit part dsigmaydp(ip) (enters sllh through dJydsigma, reference src/model.cpp:2340-2377)")
    A("    AB_HD static void dsigmaydp(const int ip, const double* p, const double* k, const double* y, double* out) {")
    A("        (void)p; (void)k; (void)y; (void)ip;")
    for iy in range(ny):
        cases = [(ip, W._pr(sp.diff(sigexpr[iy], spec.params[ip]))) for ip in range(npar)
                 if sp.diff(sigexpr[iy], spec.params[ip]) != 0]
        A(f"        out[{iy}] = {_select_chain('ip', cases)};")
    A("    }")

    # ---- events (reference generated root/stau/deltax/deltasx/z/rz/dzdx/drzdx/explicit_roots files; the symbolic
    # ---- construction follows python/sdist/amici/_symbolic/de_model.py:1560-1760) ----
    roots = [sp.sympify(ev.root).xreplace(spl_map) for ev in spec.events]
    state_dep = [any(r.has(xs) for xs in spec.states) for r in roots]
    ne_solver = sum(1 for d_ in state_dep if d_)
    if any(state_dep[ne_solver:]):
        raise ValueError("events with state-dependent roots must come first")
    A(f"    static constexpr int NE_SOLVER = {ne_solver};  // events tracked by the integrator's root finder")
    if ne:
        x_old_syms = [sp.Symbol(f"xold_{i}", real=True) for i in range(nx)]
        xdot_syms = [sp.Symbol(f"xdotnew_{i}", real=True) for i in range(nx)]
        xdot_old_syms = [sp.Symbol(f"xdotold_{i}", real=True) for i in range(nx)]
        stau_sym = sp.Symbol("stau_ip", real=True)
        for i in range(nx):
            subs[x_old_syms[i]] = sp.Symbol(f"x_old[{i}]")
            subs[xdot_syms[i]] = sp.Symbol(f"xdot[{i}]")
            subs[xdot_old_syms[i]] = sp.Symbol(f"xdot_old[{i}]")
        subs[stau_sym] = sp.Symbol("stau")
        sig_ev = "const double t, const double* x, const double* p, const double* k, const double* h, const double* spl"
        z2event = [zo.event for zo in spec.event_observables]
        zexpr = [sp.sympify(zo.formula).xreplace(spl_map) for zo in spec.event_observables]
        sigz = [sp.sympify(zo.sigma) for zo in spec.event_observables]
        drootdt_total = [sp.diff(r, t_sym) + sum((sp.diff(r, spec.states[c]) * xdot[c] for c in range(nx)), sp.Integer(0))
                         for r in roots]
        drootdt_total = [sp.simplify(e) for e in drootdt_total]

        A("    AB_HD static bool event_initial_value(const int ie) {")
        A("        return " + " : ".join([f"(ie == {ie}) ? {'true' if ev.initial_value else 'false'}"
                                            for ie, ev in enumerate(spec.events)] + ["true"]) + ";")
        A("    }")
        A("    AB_HD static int z2event(const int iz) {")
        A("        return " + " : ".join([f"(iz == {iz}) ? {ie}" for iz, ie in enumerate(z2event)] + ["-1"]) + ";")
        A("    }")
        A(f"    AB_HD static void root({sig_ev}, double* root) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl;")
        A(W.body([(f"root[{ie}]", roots[ie]) for ie in range(ne)]).rstrip("\n"))
        A("    }")
        # explicit trigger time of the time-triggered events
        A("    // trigger time of a time-triggered event (reference explicit_roots_<model>.cpp)")
        A("    AB_HD static double explicit_time(const int ie, const double* p, const double* k) {")
        A("        (void)p; (void)k;")
        for ie in range(ne_solver, ne):
            sol = sp.solve(roots[ie], t_sym)
            if len(sol) != 1:
                raise ValueError(f"event {ie}: expected exactly one trigger time, got {sol}")
            A(f"        if (ie == {ie}) return {W._pr(sol[0])};")
        A("        return INFINITY;\n    }")
        # stau: sroot / drootdt_total (minus the parameter derivative of the event time)
        A("    // stau = (drootdx*sx + drootdp_ip) / drootdt_total for event ie, parameter ip (reference stau_<model>.cpp)")
        A(f"    AB_HD static double stau(const int ie, const int ip, {sig_ev}, const double* sx) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)sx; (void)ip;")
        stau_zero = []
        for ie in range(ne):
            if drootdt_total[ie] == 0:
                stau_zero.append(True)
                continue
            lin = sum((sp.diff(roots[ie], spec.states[c]) * sx_syms[c] for c in range(nx)), sp.Integer(0))
            cases = [(ip, sp.diff(roots[ie], spec.params[ip])) for ip in range(npar) if sp.diff(roots[ie], spec.params[ip]) != 0]
            stau_zero.append(lin == 0 and not cases)
            if stau_zero[-1]:
                continue
            A(f"        if (ie == {ie}) {{")
            vals = [("const double lin", lin), ("const double den", drootdt_total[ie])]
            vals += [(f"const double v_{ip}", e) for ip, e in cases]
            A(W.body(vals, indent="            ").rstrip("\n"))
            A(f"            return (lin + ({_select_chain('ip', [(ip, f'v_{ip}') for ip, _ in cases])})) / den;")
            A("        }")
        A("        return 0.0;\n    }")
        # deltax
        deltax = []
        for ie, ev in enumerate(spec.events):
            if not ev.assignments:
                deltax.append(None)
                continue
            to_old = dict(zip(spec.states, x_old_syms))
            deltax.append([(sp.sympify(ev.assignments[xs]).xreplace(to_old) - xs) if xs in ev.assignments else sp.Integer(0)
                           for xs in spec.states])
        A("    // state update x += deltax (reference deltax_<model>.cpp); x_old = state the assignment is evaluated on")
        A("    AB_HD static void deltax(const int ie, const double t, const double* x, const double* p, const double* k, const double* h, const double* x_old, double* out) {")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)x_old; (void)ie;")
        for r in range(nx):
            A(f"        out[{r}] = 0.0;")
        for ie in range(ne):
            if deltax[ie] is None:
                continue
            A(f"        if (ie == {ie}) {{")
            A(W.body([(f"out[{r}]", deltax[ie][r]) for r in range(nx)], indent="            ").rstrip("\n"))
            A("        }")
        A("    }")
        A("    AB_HD static bool event_updates_state(const int ie) {")
        A("        return " + " || ".join([f"ie == {ie}" for ie in range(ne) if deltax[ie] is not None] + ["false"]) + ";")
        A("    }")
        # deltasx
        A("    // sensitivity update sx_ip += deltasx (reference deltasx_<model>.cpp; de_model.py:1700-1750)")
        A(f"    AB_HD static void deltasx(const int ie, const int ip, {sig_ev}, const double* xdot, const double* xdot_old, const double* sx, const double stau, const double* x_old, double* out) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)xdot; (void)xdot_old; (void)sx; (void)stau; (void)x_old; (void)ip;")
        for r in range(nx):
            A(f"        out[{r}] = 0.0;")
        for ie in range(ne):
            base = [sp.Integer(0)] * nx
            if not stau_zero[ie]:
                base = [(xdot_syms[r] - xdot_old_syms[r]) * stau_sym for r in range(nx)]
            percase = {}
            if deltax[ie] is not None:
                dx_ = deltax[ie]
                ddx = [[sp.diff(dx_[r], spec.states[c]) + sp.diff(dx_[r], x_old_syms[c]) for c in range(nx)] for r in range(nx)]
                tmp_dxdp = [sx_syms[c] - (xdot_old_syms[c] * stau_sym if not stau_zero[ie] else 0) for c in range(nx)]
                for r in range(nx):
                    add = sum((ddx[r][c] * tmp_dxdp[c] for c in range(nx)), sp.Integer(0))
                    if not stau_zero[ie]:
                        add += sp.diff(dx_[r], t_sym) * (-stau_sym)
                    base[r] = base[r] + add
                    for ip in range(npar):
                        e = sp.diff(dx_[r], spec.params[ip])
                        if e != 0:
                            percase.setdefault(r, []).append((ip, W._pr(e)))
            if all(b_ == 0 for b_ in base) and not percase:
                continue
            A(f"        if (ie == {ie}) {{")
            A(W.body([(f"const double b_{r}", base[r]) for r in range(nx)], indent="            ").rstrip("\n"))
            for r in range(nx):
                A(f"            out[{r}] = b_{r} + ({_select_chain('ip', percase.get(r, []))});")
            A("        }")
        A("    }")
        # z, sz, rz, srz, sigmaz
        dtaudx = [[(sp.diff(roots[ie], spec.states[c]) / drootdt_total[ie]) if drootdt_total[ie] != 0 else sp.Integer(0)
                   for c in range(nx)] for ie in range(ne)]
        A("    // event observables of event ie (only the entries iz with z2event(iz) == ie are written; reference z_/rz_<model>.cpp)")
        A(f"    AB_HD static void z(const int ie, {sig_ev}, double* z) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)ie; (void)z;")
        for iz in range(nz):
            A(f"        if (ie == {z2event[iz]}) z[{iz}] = {W._pr(zexpr[iz])};")
        A("    }")
        A(f"    AB_HD static void rz(const int ie, {sig_ev}, double* rz) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)ie; (void)rz;")
        for iz in range(nz):
            A(f"        if (ie == {z2event[iz]}) rz[{iz}] = {W._pr(roots[z2event[iz]])};")
        A("    }")
        A("    // sz = dzdx*sx + dzdp with dzdx including -dz/dt * dtau/dx (reference src/model.cpp:1267-1296; de_model.py:1622-1641)")
        A(f"    AB_HD static void sz(const int ie, const int ip, {sig_ev}, const double* sx, double* sz) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)ie; (void)ip; (void)sx; (void)sz;")
        for iz in range(nz):
            ie = z2event[iz]
            lin = sum(((sp.diff(zexpr[iz], spec.states[c]) - sp.diff(zexpr[iz], t_sym) * dtaudx[ie][c]) * sx_syms[c]
                       for c in range(nx)), sp.Integer(0))
            cases = [(ip, W._pr(sp.diff(zexpr[iz], spec.params[ip]))) for ip in range(npar) if sp.diff(zexpr[iz], spec.params[ip]) != 0]
            A(f"        if (ie == {ie}) sz[{iz}] = ({W._pr(sp.simplify(lin))}) + ({_select_chain('ip', cases)});")
        A("    }")
        A("    // srz = drootdx*sx + drootdp (reference src/model.cpp:1317-1350)")
        A(f"    AB_HD static void srz(const int ie, const int ip, {sig_ev}, const double* sx, double* srz) {{")
        A("        (void)t; (void)x; (void)p; (void)k; (void)h; (void)spl; (void)ie; (void)ip; (void)sx; (void)srz;")
        for iz in range(nz):
            ie = z2event[iz]
            lin = sum((sp.diff(roots[ie], spec.states[c]) * sx_syms[c] for c in range(nx)), sp.Integer(0))
            cases = [(ip, W._pr(sp.diff(roots[ie], spec.params[ip]))) for ip in range(npar) if sp.diff(roots[ie], spec.params[ip]) != 0]
            A(f"        if (ie == {ie}) srz[{iz}] = ({W._pr(lin)}) + ({_select_chain('ip', cases)});")
        A("    }")
        A("    AB_HD static void sigmaz(const double* p, const double* k, double* sigmaz) {")
        A("        (void)p; (void)k; (void)sigmaz;")
        if nz:
            A(W.body([(f"sigmaz[{iz}]", sigz[iz]) for iz in range(nz)]).rstrip("\n"))
        A("    }")
        A("    AB_HD static void dsigmazdp(const int ip, const double* p, const double* k, double* out) {")
        A("        (void)p; (void)k; (void)ip; (void)out;")
        for iz in range(nz):
            cases = [(ip, W._pr(sp.diff(sigz[iz], spec.params[ip]))) for ip in range(npar) if sp.diff(sigz[iz], spec.params[ip]) != 0]
            A(f"        out[{iz}] = {_select_chain('ip', cases)};")
        A("    }")

    if not ne:
        A("    // no events: stubs so that the event code paths compile away")
        A("    AB_HD static bool event_initial_value(const int) { return true; }")
        A("    AB_HD static int z2event(const int) { return -1; }")
        A("    AB_HD static void root(const double, const double*, const double*, const double*, const double*, const double*, double*) {}")
        A("    AB_HD static double explicit_time(const int, const double*, const double*) { return INFINITY; }")
        A("    AB_HD static double stau(const int, const int, const double, const double*, const double*, const double*, const double*, const double*, const double*) { return 0.0; }")
        A("    AB_HD static void deltax(const int, const double, const double*, const double*, const double*, const double*, const double*, double*) {}")
        A("    AB_HD static bool event_updates_state(const int) { return false; }")
        A("    AB_HD static void deltasx(const int, const int, const double, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double, const double*, double*) {}")
        A("    AB_HD static void z(const int, const double, const double*, const double*, const double*, const double*, const double*, double*) {}")
        A("    AB_HD static void rz(const int, const double, const double*, const double*, const double*, const double*, const double*, double*) {}")
        A("    AB_HD static void sz(const int, const int, const double, const double*, const double*, const double*, const double*, const double*, const double*, double*) {}")
        A("    AB_HD static void srz(const int, const int, const double, const double*, const double*, const double*, const double*, const double*, const double*, double*) {}")
        A("    AB_HD static void sigmaz(const double*, const double*, double*) {}")
        A("    AB_HD static void dsigmazdp(const int, const double*, const double*, double*) {}")

    # splines: node positions, node values and their parameter derivative (select chain)
    if nspl:
        A("    // spline node data (reference create_splines_<model>.cpp / dspline_valuesdp_<model>.cpp)")
        A("    AB_HD static double spline_node(const int ispl, const int i) {")
        for i, s in enumerate(spec.splines):
            arr = ", ".join(repr(float(v)) for v in s.node_list())
            A(f"        {{ const double n{i}[{len(s.values)}] = {{{arr}}}; if (ispl == {i}) return n{i}[i]; }}")
        A("        return 0.0;\n    }")
        BCV = {"given": -1, "zeroDerivative": 0, "natural": 1, "periodic": 3}
        EPV = {"noExtrapolation": -1, "constant": 0, "linear": 1, "polynomial": 2, "periodic": 3}
        A("    // HermiteSpline configuration per spline (reference create_splines_<model>.cpp constructor arguments)")
        A("    AB_HD static constexpr int spline_n_nodes(const int ispl) { return " +
          " : ".join(f"(ispl == {i}) ? {len(s.values)}" for i, s in enumerate(spec.splines)) + " : 2; }")
        A("    AB_HD static constexpr SplineCfg spline_cfg(const int ispl) {")
        A("        return " + " : ".join(
            f"(ispl == {i}) ? SplineCfg{{{BCV[s.bc[0]]}, {BCV[s.bc[1]]}, {EPV[s.extrapolate[0]]}, {EPV[s.extrapolate[1]]}, "
            f"{'true' if s.derivatives is None else 'false'}, {'true' if s.equidistant else 'false'}, {'true' if s.logarithmic else 'false'}}}"
            for i, s in enumerate(spec.splines)) + " : spline_cfg_default(true);")
        A("    }")
        A("    AB_HD static void spline_slopes(const int ispl, const double* p, const double* k, double* m) {")
        A("        (void)p; (void)k; (void)ispl; (void)m;")
        for i, s in enumerate(spec.splines):
            if s.derivatives is not None:
                A(f"        if (ispl == {i}) {{")
                A(W.body([(f"m[{j}]", sp.sympify(e)) for j, e in enumerate(s.derivatives)], indent="            ").rstrip("\n"))
                A("        }")
        A("    }")
        A("    AB_HD static void spline_dslopesdp(const int ispl, const int ip, const double* p, const double* k, double* dm) {")
        A("        (void)p; (void)k; (void)ip; (void)ispl; (void)dm;")
        for i, s in enumerate(spec.splines):
            if s.derivatives is not None:
                A(f"        if (ispl == {i}) {{")
                for j, e in enumerate(s.derivatives):
                    cases = [(ip, W._pr(sp.diff(sp.sympify(e), spec.params[ip]))) for ip in range(npar) if sp.diff(sp.sympify(e), spec.params[ip]) != 0]
                    A(f"            dm[{j}] = {_select_chain('ip', cases)};")
                A("        }")
        A("    }")
        A("    AB_HD static void spline_values(const int ispl, const double* p, const double* k, double* v) {")
        A("        (void)p; (void)k;")
        for i, s in enumerate(spec.splines):
            A(f"        if (ispl == {i}) {{")
            A(W.body([(f"v[{j}]", e) for j, e in enumerate(s.values)], indent="            ").rstrip("\n"))
            A("        }")
        A("    }")
        A("    AB_HD static void spline_dvaluesdp(const int ispl, const int ip, const double* p, const double* k, double* dv) {")
        A("        (void)p; (void)k; (void)ip;")
        for i, s in enumerate(spec.splines):
            A(f"        if (ispl == {i}) {{")
            for j, e in enumerate(s.values):
                cases = [(ip, W._pr(sp.diff(e, spec.params[ip]))) for ip in range(npar) if sp.diff(e, spec.params[ip]) != 0]
                A(f"            dv[{j}] = {_select_chain('ip', cases)};")
            A("        }")
        A("    }")
    else:
        A("    AB_HD static double spline_node(const int, const int) { return 0.0; }")
        A("    AB_HD static void spline_values(const int, const double*, const double*, double*) {}")
        A("    AB_HD static void spline_dvaluesdp(const int, const int, const double*, const double*, double*) {}")
    # derived constants hoisted out of xdot / jac / dxdotdp (see _FnWriter): k[NK .. NK_EXT)
    A(f"    static constexpr int NK_EXT = {nk + len(W.derived)};  // fixed parameters + derived constants")
    A("    // k[NK + i] <- parameter-only sub-expressions of the right-hand side, evaluated once per condition")
    A("    AB_HD static void derived(const double* p, double* k) {")
    A("        (void)p; (void)k;")
    for i, e in enumerate(W.derived):
        A(f"        k[{nk + i}] = {W._pr(e)};")
    A("    }")
    A("};\n")
    A("}}  // namespace amici_b200::models")
    A(f"#endif  // {guard}")
    return "\n".join(out) + "\n"


def write_model(spec: ModelSpec, out_dir: str) -> str:
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"{spec.name}.cuh")
    txt = emit_model(spec)
    if not os.path.exists(path) or open(path).read() != txt:
        with open(path, "w") as f:
            f.write(txt)
    return path
