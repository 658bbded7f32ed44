"""Engine headers of models that exist ONLY as the reference's symbolic model object (`DEModel`): built with the reference's own
classes (tests/de_models.py), emitted through amici_b200.codegen.de_export — the CUDA back end of the reference's `DEExporter`
(SURVEY.md §8(f)2).  Needs /root/reference (tests/ref_python.py); the written headers are committed, and
tests/test_cpu.py::test_de_export_only_models_are_reproduced re-derives them.

  python tests/golden/make_de_models.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from amici_b200.codegen import de_export  # noqa: E402
from tests import de_models  # noqa: E402

DE_ONLY = {"dirac": "from the reference's DEModel of model_dirac_py (tests/de_models.py::dirac) by amici_b200/codegen/de_export.py",
           "nested_events": "from the reference's DEModel of model_nested_events_py (tests/de_models.py::nested_events) by amici_b200/codegen/de_export.py",
           "calvetti": "from the reference's DEModel of model_calvetti_py, an index-1 DAE with time-triggered switches (tests/de_models.py::calvetti), by amici_b200/codegen/de_export.py",
           "robertson": "from the reference's DEModel of model_robertson_py, the index-1 DAE form (tests/de_models.py::robertson), by amici_b200/codegen/de_export.py"}


def source(name: str) -> str:
    return de_export.emit_cuda_source(de_models.build(name), name, doc=DE_ONLY[name])


if __name__ == "__main__":
    for name in DE_ONLY:
        path = os.path.join(ROOT, "amici_b200", "csrc", "models", name + ".cuh")
        with open(path, "w") as f:
            f.write(source(name))
        print("wrote", path)
