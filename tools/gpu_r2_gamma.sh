#!/bin/bash
# cfg1 companions: timing of gamma / ln|gamma| at 2^27 points and ncu --set full with source of both kernels.
tag=${1:-gm}
mkdir -p gpurun_out
python tools/bench_configs.py --reps 5 --cfg 1 | cut -c1-160
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_real_map' --launch-skip 9 --launch-count 4 \
    -o gpurun_out/full_${tag}_gamma -f python tools/bench_configs.py --reps 1 --cfg 1 > gpurun_out/ncu_${tag}_gamma.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_gamma.ncu-rep
