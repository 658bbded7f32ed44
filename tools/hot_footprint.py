#!/usr/bin/env python
"""Dynamic view of one kernel from an ncu report: executed instructions, stall samples and HOT code bytes per
source function.

  python tools/hot_footprint.py gpurun_out/prof.ncu-rep build/obj/jakstat_adjoint.o 'ELi32ELi1ELi64' [--steps S] [--lines]

Joins `ncu --page source --csv` (per-SASS-address counters) with `nvdisasm -gi` line info of the same object.
An instruction is "hot" if it was executed at least 0.2 x S times (S = number of BDF steps in the launch; default:
the executed count of the most frequent non-loop instruction is estimated by the 90th percentile of counts).
"""
from __future__ import annotations

import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_footprint import function_ranges  # noqa: E402


def sass_lines(obj, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    in_k, cur, first = False, ("?", 0), True
    out = {}
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
            if first:
                cur = (os.path.normpath(m.group(1)), int(m.group(2)))
                first = False
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
        if m:
            out[int(m.group(1), 16)] = cur
            first = True
    return out


def main():
    rep, obj, pattern = sys.argv[1:4]
    steps = None
    if "--steps" in sys.argv:
        steps = float(sys.argv[sys.argv.index("--steps") + 1])
    by_line = "--lines" in sys.argv
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = txt.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    base = int(rows[0]["Address"], 16)
    lmap = sass_lines(obj, pattern)
    counts = [int(r["Instructions Executed"]) for r in rows]
    if steps is None:
        steps = sorted(counts)[int(0.9 * len(counts))]
    thresh = 0.2 * steps
    franges = {}
    dyn, hot, stall, noinst, thr = (collections.Counter() for _ in range(5))
    tot_dyn = tot_hot = tot_stall = 0
    for r in rows:
        off = int(r["Address"], 16) - base
        f, ln = lmap.get(off, ("?", 0))
        if by_line:
            key = "%s:%d" % (os.path.basename(f), ln)
        else:
            if f not in franges:
                franges[f] = function_ranges(f)
            name = "?"
            for s, nm in franges[f]:
                if s <= ln:
                    name = nm
                else:
                    break
            key = "%s:%s" % (os.path.basename(f), name)
        n = int(r["Instructions Executed"])
        dyn[key] += n
        thr[key] += int(r.get("Thread Instructions Executed", 0) or 0)
        tot_dyn += n
        if n >= thresh:
            hot[key] += 16
            tot_hot += 16
        s = int(r["Warp Stall Sampling (All Samples)"])
        stall[key] += s
        tot_stall += s
        noinst[key] += int(r.get("stall_no_inst", 0) or 0)
    print("steps=%.3g  executed warp-instr=%.4g (%.0f per step)  hot bytes=%d  stall samples=%d" %
          (steps, tot_dyn, tot_dyn / steps, tot_hot, tot_stall))
    print("%-44s %9s %6s %8s %6s %7s %6s" % ("function", "instr/step", "dyn%", "hotB", "stall%", "noinst%", "lanes"))
    for k, n in dyn.most_common(45):
        print("%-44s %9.1f %6.1f %8d %6.1f %7.1f %6.1f" % (k, n / steps, 100.0 * n / tot_dyn, hot[k], 100.0 * stall[k] / max(1, tot_stall),
                                                        100.0 * noinst[k] / max(1, stall[k]), thr[k] / max(1, n)))


if __name__ == "__main__":
    main()
