#!/bin/bash
# ncu --set full captures of the cfg3 kernels (I call, K call).  usage: bash tools/gpu_r2_ncu.sh <tag> [n]
tag=${1:-n}; n=${2:-4194304}
mkdir -p gpurun_out
rx='k_debye|k_kpass|k_ipass|k_classify'
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 16 --launch-count 8 -o gpurun_out/full_${tag}_i -f \
    python tools/bench_configs.py --reps 1 --n $n --cfg 3 > gpurun_out/ncu_${tag}_i.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 50 --launch-count 9 -o gpurun_out/full_${tag}_k -f \
    python tools/bench_configs.py --reps 1 --n $n --cfg 3 > gpurun_out/ncu_${tag}_k.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_*.ncu-rep
