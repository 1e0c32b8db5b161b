#!/bin/bash
# Quick GPU visit + one ncu --set full capture of chosen kernels of the bench step.
# usage: bash tools/gpu_prof.sh <tag> <kernel-regex> [launch-skip] [launch-count]
tag=${1:-p}; rx=${2:-k_kpass}; skip=${3:-0}; cnt=${4:-4}
bash tools/gpu_quick.sh ${tag} -x
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip --launch-count $cnt \
    -o gpurun_out/full_${tag} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/full_${tag}.ncu-rep
