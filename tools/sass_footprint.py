#!/usr/bin/env python
"""Static code footprint of one kernel by source function (instruction-cache budget tool).

  python tools/sass_footprint.py build/obj/jakstat_adjoint.o 'ELi32ELi1ELi64' [--lines]

Disassembles the object with `nvdisasm -gi` (needs -lineinfo), attributes every SASS instruction (16 B) of the
kernel whose mangled name contains PATTERN to the innermost source line, and sums per enclosing function
(functions are found by scanning the source for AB_HD / AB_COLD / __device__ definitions).
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys
import tempfile


def function_ranges(path):
    out = []
    try:
        src = open(path).read().split("\n")
    except OSError:
        return out
    pat = re.compile(r"^\s*(?:static\s+)?(?:template\s*<[^>]*>\s*)?(?:AB_HD|AB_COLD|__device__|__global__|static AB_HD|static inline|inline)\b.*?([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    for i, line in enumerate(src, 1):
        m = pat.match(line)
        if m and not line.strip().endswith(";"):
            out.append((i, m.group(1)))
    return out


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    by_line = "--lines" in sys.argv
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    in_k = False
    cur = ("?", 0)
    pending_first = True
    counts = collections.Counter()
    total = 0
    for line in sass.split("\n"):
        if line.startswith(".text."):
            in_k = pattern in line
            continue
        if line.startswith("//---") and ".text." in line:
            in_k = False
            continue
        if not in_k:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if pending_first:  # innermost frame is listed first
                cur = (os.path.normpath(m.group(1)), int(m.group(2)))
                pending_first = False
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            counts[cur] += 16
            total += 16
            pending_first = True
    franges = {}
    agg = collections.Counter()
    for (f, ln), b in counts.items():
        if by_line:
            agg["%s:%d" % (os.path.basename(f), ln)] += b
            continue
        if f not in franges:
            franges[f] = function_ranges(f)
        name = "?"
        for start, nm in franges[f]:
            if start <= ln:
                name = nm
            else:
                break
        agg["%s:%s" % (os.path.basename(f), name)] += b
    print("total %d bytes (%d instructions)" % (total, total // 16))
    for k, b in agg.most_common(60):
        print("%7d  %5.1f%%  %s" % (b, 100.0 * b / total, k))


if __name__ == "__main__":
    main()
