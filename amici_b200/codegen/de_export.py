"""The CUDA emitter as a back end of the reference's code exporter (north_star (1)).

The reference turns a symbolic model (`DEModel`, reference python/sdist/amici/_symbolic/de_model.py:79) into C++ with
`DEExporter` (reference python/sdist/amici/exporters/sundials/de_export.py:127; `_generate_c_code` :267-314 writes one
`<fn>.cpp` per model function).  This module lets the same `DEModel` object drive the CUDA emitter:

* `modelspec_from_de_model(de_model, name)` reads the model definition back out of a `DEModel` through its public accessors
  (`differential_states()`, `free_parameters()`, `fixed_parameters()`, `expressions()`, `observables()`, `sigma_ys()`,
  `events()`, `event_observables()`, `sigma_zs()`, `log_likelihood_ys()`, the spline list) and builds the emitter's `ModelSpec`;
* `emit_cuda(de_model, out_dir, name)` writes `<name>.cuh`, the header a translation unit `csrc/tu/<name>.cu` instantiates;
* `install_hook(DEExporter)` wraps `DEExporter._generate_c_code` so that every export also writes `<model_name>.cuh` into the
  model directory — the three lines a maintainer would add to `_generate_c_code` itself.

The module is duck-typed: it imports nothing from the reference, so it also runs where the reference's package is not
installed (the objects only have to offer the accessors named above).  What the engine does not implement is refused with
`CudaEmitterError` instead of being emitted wrongly: DAE models with state-changing / state-triggered events, conservation laws, observable
transformations other than `lin`, noise models other than the Gaussian one, event priorities.

Tests: tests/test_cpu.py::test_de_export_* build the reference's `DEModel` for the shipped test models with the reference's own
classes (tests/ref_python.py makes the pure-Python part of the reference importable here) and require the emitted header to be
IDENTICAL to the committed `csrc/models/<name>.cuh` — so every parity test of those models is a test of this path — and the
emitter's derivatives to equal the ones `DEModel` derives (`eq('dxdotdp_explicit')`, `eq('dydx')`, ...).
"""
from __future__ import annotations

import os

import sympy as sp

from . import emitter as em


class CudaEmitterError(NotImplementedError):
    pass


def _retime(expr, de_model):
    """the reference's time symbol (`amici_time_symbol`, a real Symbol named t) -> the emitter's"""
    expr = sp.sympify(expr)
    for s in expr.free_symbols:
        if s.name == "t" and s != em.t_sym:
            expr = expr.subs(s, em.t_sym)
    return expr


def _inline_expressions(de_model):
    """{w symbol: its formula with every other w substituted} (the emitter inlines `w`, emitter.py docstring)"""
    # a spline is an expression whose value is the opaque function AmiciSpline(id, t, *parameters) (reference
    # python/sdist/amici/importers/sbml/splines.py:1374-1450, DEModel.add_spline de_model.py:540-556): its symbol stays, the
    # emitter evaluates it from the spline table (emitter.Spline)
    spline_ids = {s.sbml_id for s in getattr(de_model, "_splines", [])}
    exprs = [e for e in de_model.expressions() if e.get_sym() not in spline_ids]
    table = {e.get_sym(): sp.sympify(e.get_val()) for e in exprs}
    done: dict = {}
    pending = dict(table)
    for _ in range(len(pending) + 1):
        progressed = False
        for s, v in list(pending.items()):
            if not (v.free_symbols & (set(pending) - {s})):
                done[s] = v.subs(done) if done else v
                del pending[s]
                progressed = True
        if not pending:
            break
        if not progressed:
            raise CudaEmitterError("circular dependencies among the model expressions: %s" % sorted(map(str, pending)))
    return done


def _check_gaussian(de_model, obs, sigmas):
    """`Jy` must be the negative log-likelihood of a Gaussian with the observable's sigma (the engine's `Jy`, simulate.cuh)"""
    llhs = list(de_model.log_likelihood_ys())
    if not llhs:
        return
    if len(llhs) != len(obs):
        raise CudaEmitterError("one log-likelihood per observable expected")
    for o, sg, ll in zip(obs, sigmas, llhs):
        y, my, s = o.get_sym(), o.get_measurement_symbol(), sg.get_sym()
        gauss = sp.Rational(1, 2) * sp.log(2 * sp.pi * s ** 2) + sp.Rational(1, 2) * ((y - my) / s) ** 2
        diff = sp.simplify(sp.nsimplify(sp.sympify(ll.get_val()) - gauss, rational=True))
        if diff != 0:
            raise CudaEmitterError("noise model of observable %s is not Gaussian: %s" % (y, ll.get_val()))


def _spline(spl, inline, de_model):
    """reference python/sdist/amici/importers/sbml/splines.py CubicHermiteSpline -> emitter.Spline"""
    def ex(v):
        return _retime(sp.sympify(v).subs(inline), de_model)

    bc = tuple("zeroDerivative" if b == "zeroderivative" else (b or "given") for b in getattr(spl, "bc", (None, None)))
    extr = tuple({None: "noExtrapolation", "constant": "constant", "linear": "linear", "polynomial": "polynomial",
                  "periodic": "periodic"}[e] for e in getattr(spl, "extrapolate", (None, None)))
    nodes = spl.nodes
    equidistant = hasattr(nodes, "start")     # UniformGrid
    node_vals = [float(nodes.start), float(nodes.stop)] if equidistant else [float(sp.sympify(v)) for v in nodes]
    derivs = None if getattr(spl, "derivatives_by_fd", True) else [ex(v) for v in spl.derivatives_at_nodes]
    return em.Spline(str(spl.sbml_id), node_vals, [ex(v) for v in spl.values_at_nodes],
                     logarithmic=bool(getattr(spl, "logarithmic_parametrization", False)), derivatives=derivs, bc=bc,
                     extrapolate=extr, equidistant=equidistant)


def modelspec_from_de_model(de_model, name: str, doc: str = "") -> em.ModelSpec:
    alg_states, alg_eqs = list(de_model.algebraic_states()), list(de_model.algebraic_equations())
    if len(alg_states) != len(alg_eqs):
        raise CudaEmitterError("one algebraic equation per algebraic state expected")
    if alg_states:
        # the DAE path (csrc/dae.cuh) handles time-triggered Heaviside switches only
        state_syms = {s.get_sym() for s in de_model.differential_states()} | {s.get_sym() for s in alg_states}
        for e in de_model.events():
            if e._assignments or (sp.sympify(e.get_val()).subs(_inline_expressions(de_model)).free_symbols & state_syms):
                raise CudaEmitterError("DAE models with events that change the state or are triggered by it are not supported (csrc/dae.cuh)")
        if list(de_model.event_observables()) or getattr(de_model, "_splines", []):
            raise CudaEmitterError("DAE models with event observables or splines are not supported (csrc/dae.cuh)")
    if de_model.num_cons_law() > 0 or list(de_model.conservation_laws()):
        raise CudaEmitterError("conservation laws (x_rdata != x_solver) are not supported: import with compute_conservation_laws=False")
    inline = _inline_expressions(de_model)

    def ex(v):
        return _retime(sp.sympify(v).subs(inline), de_model)

    xs = list(de_model.differential_states())
    states = [s.get_sym() for s in xs] + [s.get_sym() for s in alg_states]
    ps, ks = list(de_model.free_parameters()), list(de_model.fixed_parameters())
    obs, sigmas = list(de_model.observables()), list(de_model.sigma_ys())
    if len(sigmas) != len(obs):
        raise CudaEmitterError("one sigma per observable expected")
    for o in obs:
        if getattr(o, "trafo", "lin") not in (None, "lin"):
            raise CudaEmitterError("observable transformation %r of %s is not supported" % (o.trafo, o.get_sym()))
    _check_gaussian(de_model, obs, sigmas)
    sig_subs = {sg.get_sym(): sp.sympify(sg.get_val()) for sg in sigmas}
    observables = [em.Observable(str(o.get_sym()), ex(o.get_val()), ex(sp.sympify(sg.get_val()).subs(sig_subs)))
                   for o, sg in zip(obs, sigmas)]

    evs = list(de_model.events())
    ev_index = {e.get_sym(): i for i, e in enumerate(evs)}
    events = []
    for e in evs:
        if e.get_priority() is not None:
            raise CudaEmitterError("event priorities are not supported")
        if e._assignments and not e._use_values_from_trigger_time:
            pass   # simultaneous assignments are evaluated on the pre-event state either way (emitter.Event docstring)
        assignments = {s: ex(v) for s, v in e._assignments.items()} or None
        events.append(em.Event(ex(e.get_val()), assignments=assignments, initial_value=bool(e.get_initial_value())))
    used = set().union(*[ex(s.get_dt()).free_symbols for s in xs], *[o.formula.free_symbols for o in observables]) if xs else set()
    heavisides = {s: i for s, i in ev_index.items() if s in used} or None

    zs, sigma_zs = list(de_model.event_observables()), list(de_model.sigma_zs())
    event_observables = []
    for i, z in enumerate(zs):
        ev = z.get_event()
        if ev not in ev_index:
            raise CudaEmitterError("event observable %s refers to an unknown event %s" % (z.get_sym(), ev))
        sg = ex(sigma_zs[i].get_val()) if i < len(sigma_zs) else sp.Float(1.0)
        event_observables.append(em.EventObservable(str(z.get_sym()), ev_index[ev], ex(z.get_val()), sg))

    splines = [_spline(s, inline, de_model) for s in getattr(de_model, "_splines", [])]
    return em.ModelSpec(
        name=name, states=states, params=[p.get_sym() for p in ps], fixed=[k.get_sym() for k in ks],
        xdot=[ex(s.get_dt()) for s in xs] + [ex(e.get_val()) for e in alg_eqs],
        x0=[ex(s.get_val()) for s in xs] + [ex(s.get_val()) for s in alg_states], observables=observables,
        algebraic=[False] * len(xs) + [True] * len(alg_states) if alg_states else (),
        p_nominal=[float(p.get_val()) for p in ps], k_nominal=[float(k.get_val()) for k in ks], splines=splines,
        events=events, event_observables=event_observables, heavisides=heavisides, doc=doc)


def emit_cuda_source(de_model, name: str, doc: str = "") -> str:
    return em.emit_model(modelspec_from_de_model(de_model, name, doc))


def emit_cuda(de_model, out_dir: str, name: str, doc: str = "") -> str:
    """-> path of the written `<name>.cuh`"""
    return em.write_model(modelspec_from_de_model(de_model, name, doc), out_dir)


def install_hook(exporter_cls):
    """Make `exporter_cls` (the reference's `DEExporter`) write `<model_name>.cuh` next to the C++ sources on every export."""
    if getattr(exporter_cls, "_amici_b200_hook", False):
        return exporter_cls
    original = exporter_cls._generate_c_code

    def _generate_c_code(self):
        original(self)
        emit_cuda(self.model, os.fspath(self.model_path), self.model_name)

    exporter_cls._generate_c_code = _generate_c_code
    exporter_cls._amici_b200_hook = True
    return exporter_cls
