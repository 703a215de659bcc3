#!/bin/bash
# first contact with the GPU: the gpu test-suite without -x so one call shows every failing stage
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
