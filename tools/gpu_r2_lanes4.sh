#!/bin/bash
# duration / active lanes of every kernel of the cfg4 calls (H1, H2, Ai, Bi).  usage: bash tools/gpu_r2_lanes4.sh <tag>
tag=${1:-l}
ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:'k_' --csv --log-file gpurun_out/lanes4_${tag}.csv python tools/bench_configs.py --reps 1 --n 4194304 --cfg 4 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/lanes4_${tag}.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: h=r; rows=rows[i+1:]; break
k,m,v=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value")
seen=collections.OrderedDict()
for r in rows:
    if len(r)>v: seen.setdefault((r[0],r[k][:60]),{})[r[m].split("__")[1][:14]]=r[v]
last=collections.OrderedDict()
for (i,n),d in seen.items(): last[n]=d
for n,d in last.items(): print("%-62s %s" % (n, d))
PY
