#!/bin/bash
# A/B of library variants on one box: bash tools/gpu_variants.sh <tag> exp/lib_a.so exp/lib_b.so ...
# (each variant replaces the in-tree library for one short bench + the all-config table)
tag=$1; shift
cp sciport_rs_b200/libsciport_b200.so /tmp/lib_base.so
for lib in base "$@"; do
  [ "$lib" != base ] && cp "$lib" sciport_rs_b200/libsciport_b200.so
  name=$(basename "$lib" .so)
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/var_${tag}_${name}.json 2> gpurun_out/var_${tag}_${name}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/var_${tag}_${name}.json").read().strip().splitlines()[-1])
print("${name}: ms/step %.4f" % d["ms_per_step"], " ".join("%s=%.3f" % (k["name"].split("/")[-1], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.02), "err %.2e" % d["accuracy"]["max_norm_err"])
PY
  if [ -n "$VAR_CONFIGS" ]; then
    python tools/bench_configs.py --cfg $VAR_CONFIGS 2> /dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   cfg%d %-26s %8.3f ms' % (d['cfg'], d['fn'], d['ms']), ' '.join('%s=%.3f' % (k['name'].split('/')[-1], k['ms']) for k in d.get('kernels', []) if k['ms'] > 0.05))
"
  fi
done
cp /tmp/lib_base.so sciport_rs_b200/libsciport_b200.so
