#!/bin/bash
# the bench command on its own (must exit 0), then its ncu launch list only (gpu__time_duration.sum, --clock-control none)
tag=${1:-ll}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-configs"
$B > gpurun_out/bench_${tag}_plain.json 2> gpurun_out/bench_${tag}_plain.err
rc=$?; echo "plain bench rc=$rc"; [ $rc -ne 0 ] && exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_${tag}.csv \
    $B > gpurun_out/ncu_${tag}_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_${tag}.csv
