#!/bin/bash
# ncu --set full captures of the heavy cfg3 kernels (I call: classify, Miller, Wronskian, Debye; K call: classify,
# Debye, K pass).  Reports must stay below the 64 MiB that gpurun copies back.  usage: bash tools/gpu_r2_ncu.sh <tag> [n]
tag=${1:-n}; n=${2:-4194304}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_debye|k_ipass_sorted|k_ipass<2|k_classify' --launch-skip 8 --launch-count 4 \
    -o gpurun_out/full_${tag}_i -f python tools/bench_configs.py --reps 1 --n $n --cfg 3 > gpurun_out/ncu_${tag}_i.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_debye|k_kpass|k_classify' --launch-skip 14 --launch-count 3 \
    -o gpurun_out/full_${tag}_k -f python tools/bench_configs.py --reps 1 --n $n --cfg 3 > gpurun_out/ncu_${tag}_k.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_*.ncu-rep; du -sm gpurun_out
