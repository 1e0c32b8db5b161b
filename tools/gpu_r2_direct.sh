#!/bin/bash
# k_direct visit: its parity tests, then cfg2 / cfg4 / cfg1-sized timings with the kernel on and off (SPB_DIRECT)
tag=${1:-d1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "direct or scalar_negative or fused_large or compaction or golden_device or reference_surface" > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_${tag}.log
bash tools/gpu_r2_ab.sh ${tag} "SPB_DIRECT=1" "SPB_DIRECT=0" "SPB_DIRECT=1" "SPB_DIRECT=0"
