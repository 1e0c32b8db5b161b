#!/bin/bash
# ncu --set full with source: k_direct<JY> (SPB_DIRECT=1) against the list-driven fused kernel (SPB_DIRECT=0) on cfg2
tag=${1:-dn}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_direct<\(int\)6' --launch-skip 3 --launch-count 1 \
    -o gpurun_out/full_${tag}_direct -f python tools/bench_configs.py --reps 1 --cfg 2 > gpurun_out/ncu_${tag}_direct.log 2>&1
echo "rc=$?"
SPB_DIRECT=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_ipass<\(int\)1, \(bool\)1, \(int\)6' --launch-skip 3 --launch-count 1 \
    -o gpurun_out/full_${tag}_listed -f python tools/bench_configs.py --reps 1 --cfg 2 > gpurun_out/ncu_${tag}_listed.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_*.ncu-rep
