#!/bin/bash
# active lanes per instruction + duration of every kernel of one cfg3 I call and one K call.  usage: bash tools/gpu_r2_lanes.sh <tag>
tag=${1:-l}
ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:'k_' --csv --log-file gpurun_out/lanes_${tag}.csv python tools/bench_configs.py --reps 1 --n 4194304 --cfg 3 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/lanes_${tag}.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: h=r; rows=rows[i+1:]; break
k,m,v=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value")
seen=collections.OrderedDict()
for r in rows:
    if len(r)>v: seen.setdefault((r[0],r[k][:52]),{})[r[m].split("__")[1][:14]]=r[v]
items=list(seen.items())
# third call of I (warm) and third call of K: print each distinct kernel once, the last occurrence
last={}
for (i,n),d in items: last[n]=d
for n,d in last.items(): print("%-54s %s" % (n, d))
PY
