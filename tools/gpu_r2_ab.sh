#!/bin/bash
# A/B of the round-2 grouping knobs on the cfg3 calls (ive, kve, pair; 2^26 points = 4 chunks).
# usage: bash tools/gpu_r2_ab.sh <tag> ["VAR=val VAR=val" ...]   (each quoted argument is one variant's environment)
tag=${1:-ab}; shift
mkdir -p gpurun_out
i=0
for env in "$@"; do
    echo "=== variant $i: $env"
    env $env python tools/bench_configs.py --reps 5 --n ${ABN:-67108864} --cfg ${CFGS:-3} > gpurun_out/ab_${tag}_$i.jsonl 2> gpurun_out/ab_${tag}_$i.err || tail -3 gpurun_out/ab_${tag}_$i.err
    python - <<PY
import json
for l in open("gpurun_out/ab_${tag}_$i.jsonl"):
    d = json.loads(l)
    ks = d.get("kernels") or []
    agg = {}
    for k in ks:
        a = agg.setdefault(k["name"], [0.0, 0]); a[0] += k["ms"]; a[1] += k["points"]
    print("%-22s %8.3f ms  %.3e evals/s   " % (d["fn"], d["ms"], d["evals_per_s"]) + "  ".join("%s %.2f" % (n.split("/")[-1], v[0]) for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:9]))
PY
    i=$((i+1))
done
