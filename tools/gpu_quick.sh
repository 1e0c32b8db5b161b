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
