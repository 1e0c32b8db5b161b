#!/bin/bash
# cfg5 (order sweep) visit: sequence tests, timing at 2^21 points, ncu --set full of k_sweep with source.
tag=${1:-s}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "sweep or sequence or seq" 2>&1 | tail -3
python tools/bench_configs.py --reps 5 --n5 2097152 --cfg 5 | cut -c1-200
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_sweep' --launch-skip 1 --launch-count 1 \
    -o gpurun_out/full_${tag}_sweep -f python tools/bench_configs.py --reps 1 --n5 524288 --cfg 5 > gpurun_out/ncu_${tag}_sweep.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_sweep.ncu-rep
