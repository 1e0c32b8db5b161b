#!/bin/bash
# Round-2 profile visit (one GPU): the bench command on its own first (must exit 0), then
#   (1) the ncu launch list of the same command (gpu__time_duration.sum, --clock-control none),
#   (2) ncu --set full of every pipeline kernel of one steady-state chunk (no source: small report),
#   (3) ncu --set full --import-source on of the three dominant kernels (Debye-direct, Wronskian route, K pass).
# usage: bash tools/gpu_r2_profile.sh <tag>
tag=${1:-p}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-configs"
$B > gpurun_out/bench_${tag}_plain.json 2> gpurun_out/bench_${tag}_plain.err
rc=$?; echo "plain bench rc=$rc"
[ $rc -ne 0 ] && { tail -5 gpurun_out/bench_${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_${tag}.csv \
    $B > gpurun_out/ncu_${tag}_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_${tag}.csv
# one step = 16 chunks; skip the first two steps' worth of pipeline launches (cold caches, scratch growth)
K='k_debye|k_ipass|k_kpass|k_classify|k_scan|k_scatter|k_cell|k_general'
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$K" --launch-skip 480 --launch-count 16 \
    -o gpurun_out/full_${tag}_chunk -f $B > gpurun_out/ncu_${tag}_chunk.log 2>&1
echo "chunk capture rc=$?"
# the report of 16 large kernels exceeds what gpurun copies back: keep its raw page as CSV, drop the report
ncu -i gpurun_out/full_${tag}_chunk.ncu-rep --page raw --csv > gpurun_out/full_${tag}_chunk_raw.csv 2> /dev/null
rm -f gpurun_out/full_${tag}_chunk.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_debye|k_kpass|k_classify|k_ipass<\(int\)3,' \
    --launch-skip 96 --launch-count 4 -o gpurun_out/full_${tag}_top -f $B > gpurun_out/ncu_${tag}_top.log 2>&1
echo "top capture rc=$?"
ls -la gpurun_out/full_${tag}_*.ncu-rep; du -sm gpurun_out
