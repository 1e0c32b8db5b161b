#!/bin/bash
# All-config throughput + lane utilisation per kernel.  usage: bash tools/gpu_configs.sh <tag> [bench_configs args]
tag=${1:-c}
shift
mkdir -p gpurun_out
python tools/bench_configs.py "$@" > gpurun_out/configs_${tag}.jsonl 2> gpurun_out/configs_${tag}.err
echo "rc=$?"; tail -3 gpurun_out/configs_${tag}.err
python - <<PY
import json
for l in open("gpurun_out/configs_${tag}.jsonl"):
    d=json.loads(l)
    print("cfg%d %-28s %8.3f ms %10.3e evals/s %s" % (d["cfg"], d["fn"], d["ms"], d["evals_per_s"], ("%.0f GB/s" % d["gbs"]) if "gbs" in d else ""))
    for k in d.get("kernels", []):
        if k["ms"] > 0.02: print("       %-16s %8.3f ms %10d pts" % (k["name"].split("/")[-1], k["ms"], k["points"]))
PY
ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:'k_kpass|k_ipass|k_general|k_sequence|k_airy' --csv --log-file gpurun_out/lanes_cfg_${tag}.csv \
    python tools/bench_configs.py --reps 1 --n 4194304 --n5 262144 --cfg 3,4,5 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/lanes_cfg_${tag}.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: h=r; rows=rows[i+1:]; break
k,m,v=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value")
seen=collections.OrderedDict()
for r in rows:
    if len(r)>v: seen.setdefault((r[0],r[k][:34]),{})[r[m].split("__")[1][:12]]=r[v]
last=None
for (i,n),d in seen.items():
    key=(n,d.get("inst_execute"))
    if key!=last: print(n, d)
    last=key
PY
