#!/bin/bash
# usage: tools/gpu_test.sh [pytest -k expression]
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
if [ -n "$1" ]; then timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x -k "$1" 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
else timeout 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; fi
tail -40 gpurun_out/pytest_gpu.log
