#!/bin/bash
# One GPU-box visit (round 2): parity tests, smoke, bench (both arms).  usage: bash tools/gpu_r2_check.sh <tag> [pytest-args]
tag=${1:-r2}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8 > gpurun_out/box_${tag}.txt
nproc >> gpurun_out/box_${tag}.txt; free -g | head -2 >> gpurun_out/box_${tag}.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 "$@" > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_${tag}.log
tail -15 gpurun_out/pytest_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}.err
python tools/show_bench.py gpurun_out/bench_${tag}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_${tag}.json
