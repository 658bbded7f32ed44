#!/bin/bash
# usage (on the GPU box): tools/ab_variants.sh "" _s1 _h1   -> solves/s of each library variant (kernel-only bench)
for v in "$@"; do
  echo -n "variant=$v "
  AMICI_B200_LIB=$PWD/amici_b200/lib/libamici_b200$v.so timeout 150 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | python -c "import sys,json; [print(round(d[\"value\"]), round(d[\"ms_per_step\"],1), d[\"config\"][\"n_failed\"], d[\"config\"][\"sum_llh\"]) for d in [json.loads(l) for l in sys.stdin if l.startswith(\"{\")]]"
done
