#!/bin/bash
# usage (on the GPU box): tools/ncu_stalls.sh <out tag> [bench env assignments...] -- lib variants ...
# Collects the issue-stall breakdown of the solve kernel for each library variant (4736-condition launch) into gpurun_out/stalls_<tag><variant>.csv
tag=$1; shift
R=smsp__average_warps_issue_stalled
M="gpu__time_duration.sum,launch__registers_per_thread,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_per_inst_issued.ratio"
for s in no_instruction wait short_scoreboard long_scoreboard branch_resolving sleeping math_pipe_throttle barrier dispatch_stall drain imc_miss lg_throttle membar mio_throttle misc not_selected selected tex_throttle; do M="$M,${R}_${s}_per_issue_active.ratio"; done
mkdir -p gpurun_out
for v in "$@"; do
  AMICI_B200_LIB=$PWD/amici_b200/lib/libamici_b200$v.so timeout 300 ncu --metrics $M --clock-control none -k regex:bdf_ -s 1 -c 1 --csv --log-file gpurun_out/stalls_$tag$v.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-conditions 4736 > gpurun_out/stalls_$tag$v.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/stalls_$tag$v.csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value")
print("variant=$v", " ".join("%s=%s"%(r[i_n].replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","").replace(".avg.pct_of_peak_sustained_active","").replace(".ratio","").replace(".sum",""), r[i_v]) for r in rows[1:]))
PY
done
