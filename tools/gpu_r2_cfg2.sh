#!/bin/bash
# cfg2 (jv+yv pair, scalar order, 4096^2 grid): ncu --set full with source of the fused large-|w| kernel and the classifier
tag=${1:-c2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_ipass<\(int\)1, \(bool\)1, \(int\)6|k_classify<\(int\)6' --launch-skip 4 --launch-count 2 \
    -o gpurun_out/full_${tag}_cfg2 -f python tools/bench_configs.py --reps 1 --cfg 2 > gpurun_out/ncu_${tag}_cfg2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_cfg2.ncu-rep
