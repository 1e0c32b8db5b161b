#!/bin/bash
# N-GPU visit: the multi-device test (must run, not skip) + the sharded bench under torchrun.  usage: bash tools/gpu_r2_multi.sh <tag> <N>
tag=${1:-m}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader > gpurun_out/box_${tag}.txt; nproc >> gpurun_out/box_${tag}.txt
nvidia-smi topo -m >> gpurun_out/box_${tag}.txt 2>&1
python -m pytest tests -m gpu -q -x -k "multi_device" -rs 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 \
    > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${tag}.err | cut -c1-300
python tools/show_bench.py gpurun_out/bench_${tag}.json
