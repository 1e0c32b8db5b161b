#!/bin/bash
# ncu --set full of the memory-bound companions: real gamma / ln|gamma| maps (2^27 points) and the order sweep (cfg5).
tag=${1:-n}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_real_map<\(int\)[01]>|k_sweep' --launch-skip 8 --launch-count 4 \
    -o gpurun_out/full_${tag}_mem -f python tools/bench_configs.py --reps 1 --cfg 1,5 > gpurun_out/ncu_${tag}_mem.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_mem.ncu-rep; tail -3 gpurun_out/ncu_${tag}_mem.log | cut -c1-300
