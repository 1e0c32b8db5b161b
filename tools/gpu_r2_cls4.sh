#!/bin/bash
# ncu --set full with source of the classifier on config 4 (hankel1, unscaled, scalar order)
tag=${1:-c4}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_classify<\(int\)4' --launch-skip 2 --launch-count 1 \
    -o gpurun_out/full_${tag}_cls4 -f python tools/bench_configs.py --reps 1 --n 4194304 --cfg 4 > gpurun_out/ncu_${tag}_cls4.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_cls4.ncu-rep
