#!/bin/bash
# ncu --set full of the heavy kernels of the cfg3 PAIR call (Wronskian route, K pass, Miller, Debye).  usage: bash tools/gpu_r2_ncu2.sh <tag>
tag=${1:-n}
mkdir -p gpurun_out
# launches per bessel call matching the regex: I call 3 (mlri, wrsk, debye), K call 4 (+ kpass), pair 4; 4 I calls + 4 K calls = 28, third pair call starts at 28 + 8
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_debye|k_kpass|k_ipass<\(int\)[23],' --launch-skip 36 --launch-count 4 \
    -o gpurun_out/full_${tag}_ik -f python tools/bench_configs.py --reps 1 --n 4194304 --cfg 3 > gpurun_out/ncu_${tag}_ik.log 2>&1
echo "rc=$?"; ls -la gpurun_out/full_${tag}_ik.ncu-rep
