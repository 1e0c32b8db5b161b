#!/bin/bash
# ncu --set full with source of the fused large-|w| kernel of the cfg3 pair call (per-element orders)
tag=${1:-a}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_ipass<\(int\)1, \(bool\)0, \(int\)7' --launch-skip 2 --launch-count 1 \
    -o gpurun_out/full_${tag}_asyi -f python tools/bench_configs.py --reps 1 --n 16777216 --cfg 3 > gpurun_out/ncu_${tag}_asyi.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_asyi.ncu-rep
