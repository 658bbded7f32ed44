"""Test infrastructure: import the reference's PYTHON package (symbolic model + code exporter) in this container.

The reference's `amici` package cannot be installed here (no SWIG, no libsbml), but its symbolic layer —
`amici._symbolic.de_model.DEModel` and `amici.exporters.sundials.de_export.DEExporter` — is pure Python + SymPy.  This module
makes `import amici` work from `/root/reference/python/sdist` by stubbing the third-party modules the package imports at load
time and never needs for a hand-built `DEModel` (`libsbml`, `sbmlmath`, `swig`), and by providing `toposort` (Kahn's algorithm,
the published behaviour of toposort 1.x: an iterator over sets of mutually independent items).

Only `tests/` and `tests/golden/make_*.py` import this; nothing on the GPU box does (the reference is not there).
"""
import importlib.machinery
import os
import sys
import types

REF_SDIST = "/root/reference/python/sdist"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SDIST, "amici", "_symbolic"))


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (object,), {})


def _toposort(data):
    """toposort(data: dict[item, set[item]]) -> yields sets of items whose dependencies have all been yielded before"""
    data = {k: set(v) - {k} for k, v in data.items()}
    for extra in set().union(*data.values()) - set(data) if data else ():
        data[extra] = set()
    while True:
        ready = {k for k, v in data.items() if not v}
        if not ready:
            break
        yield ready
        data = {k: v - ready for k, v in data.items() if k not in ready}
    if data:
        raise ValueError("circular dependencies among %r" % sorted(map(str, data)))


def load():
    """-> the reference's `amici` package (imported once)"""
    if "amici" in sys.modules and getattr(sys.modules["amici"], "__file__", "").startswith(REF_SDIST):
        return sys.modules["amici"]
    if not available():
        raise ImportError("the reference tree is not present")
    for name in ("libsbml", "sbmlmath", "swig"):
        if name not in sys.modules:
            m = _Anything(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    if "toposort" not in sys.modules:
        m = types.ModuleType("toposort")
        m.__spec__ = importlib.machinery.ModuleSpec("toposort", None)
        m.toposort = _toposort
        m.toposort_flatten = lambda data, sort=True: [x for group in _toposort(data) for x in (sorted(group) if sort else group)]
        m.CircularDependencyError = ValueError
        sys.modules["toposort"] = m
    if REF_SDIST not in sys.path:
        sys.path.insert(0, REF_SDIST)
    import amici  # noqa: E402
    return amici
