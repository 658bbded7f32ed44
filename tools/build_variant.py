#!/usr/bin/env python
"""Build an experimental variant of the engine: the JAK-STAT translation unit recompiled with extra -D flags, linked
with the regular objects into amici_b200/lib/libamici_b200_<tag>.so (select it with AMICI_B200_LIB=<path>).

  python tools/build_variant.py nosync -DAB_STEP_SYNC=0 -DAB_THREADS_G32=64
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
tu = "jakstat_adjoint"
if flags and flags[0].startswith("--tu="):
    tu, flags = flags[0][5:], flags[1:]
ge.build_cuda()
objdir = os.path.join(ROOT, "build", "obj_" + tag)
os.makedirs(objdir, exist_ok=True)
src = os.path.join(ge.CSRC, "tu", tu + ".cu")
obj = os.path.join(objdir, tu + ".o")
ge._run(["nvcc"] + ge.NVCC_FLAGS + flags + ["-c", src, "-o", obj])
objs = [os.path.join(ge.OBJDIR, f) for f in sorted(os.listdir(ge.OBJDIR)) if f.endswith(".o") and f != tu + ".o"] + [obj]
lib = os.path.join(ge.LIBDIR, "libamici_b200_%s.so" % tag)
ge._run(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
