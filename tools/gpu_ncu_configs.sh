#!/bin/bash
# ncu --set full capture of the non-headline kernels (configs 3, 4, 5).  usage: bash tools/gpu_ncu_configs.sh <tag> <kernel-regex> <cfg-list>
tag=${1:-n}; rx=${2:-k_sweep}; cfg=${3:-5}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 2 --launch-count 6 -o gpurun_out/full_${tag} -f \
    python tools/bench_configs.py --reps 1 --n 4194304 --n5 1048576 --cfg $cfg > gpurun_out/ncu_full_${tag}.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}.ncu-rep
