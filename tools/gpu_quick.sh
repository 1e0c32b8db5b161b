#!/bin/bash
# Quick GPU visit: parity tests + bench only.  usage: bash tools/gpu_quick.sh <tag> [pytest-args]
tag=${1:-q}
shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_${tag}.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f e2e %.3e frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"] or 0))
for k in d["roofline"]["kernels"]: print("  %-18s %.3f ms  %d" % (k["name"], k["ms"], k["points"]))
PY
if [ -n "$NCU_LANES" ]; then
ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:'k_kpass|k_ipass|k_classify|k_scatter|k_general|k_sequence' --launch-skip 24 --launch-count 16 --csv --log-file gpurun_out/lanes_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/lanes_${tag}.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: h=r; rows=rows[i+1:]; break
k,m,v=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value")
seen={}
for r in rows:
    if len(r)>v: seen.setdefault((r[0],r[k][:40]),{})[r[m]]=r[v]
for (i,n),d in list(seen.items())[:16]: print(n, d)
PY
fi
