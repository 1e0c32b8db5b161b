#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and one full capture of the hot kernels.
# usage (from the repo root on the box): bash tools/gpu_check.sh <tag> [pytest-args]
tag=${1:-run}
shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x "$@" > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_${tag}.log
tail -5 gpurun_out/pytest_${tag}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_${tag}.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_${tag}.log 2>&1
echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_kpass|k_ipass|k_classify|k_scatter|k_scan' \
    --launch-skip 30 --launch-count 10 -o gpurun_out/full_${tag} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
