#!/bin/bash
# A/B of whole library builds on one box: every variants/<name>.so in turn takes the place of the in-tree library
# (on the box's scratch copy only) and runs tools/bench_configs.py; "head" = the in-tree build itself.
# usage: [CFGS=3,4] [ABN=67108864] bash tools/gpu_r2_libab.sh <tag> head prev ...
tag=${1:-lab}; shift
mkdir -p gpurun_out
cp sciport_rs_b200/libsciport_b200.so /tmp/lib_head.so
for name in "$@"; do
    if [ "$name" = head ]; then cp /tmp/lib_head.so sciport_rs_b200/libsciport_b200.so; else cp variants/lib_$name.so sciport_rs_b200/libsciport_b200.so; fi
    echo "=== $name"
    python tools/bench_configs.py --reps ${REPS:-5} --n ${ABN:-67108864} --cfg ${CFGS:-3} > gpurun_out/lab_${tag}_$name.jsonl 2> gpurun_out/lab_${tag}_$name.err || tail -3 gpurun_out/lab_${tag}_$name.err
    python - <<PY
import json
for l in open("gpurun_out/lab_${tag}_$name.jsonl"):
    d = json.loads(l)
    ks = d.get("kernels") or []
    agg = {}
    for k in ks:
        a = agg.setdefault(k["name"], [0.0, 0]); a[0] += k["ms"]; a[1] += k["points"]
    print("%-28s %8.3f ms  %.3e evals/s   " % (d["fn"][:28], d["ms"], d["evals_per_s"]) + "  ".join("%s %.3f" % (n.split("/")[-1], v[0]) for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:10]))
PY
done
cp /tmp/lib_head.so sciport_rs_b200/libsciport_b200.so
