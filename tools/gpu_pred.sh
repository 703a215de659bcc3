#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
if [ "$3" != "notest" ]; then timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log; fi
timeout 600 python tools/pred_bench.py ${1:-4} > gpurun_out/pred_bands1.log 2>&1; grep "${2:-way 0}" gpurun_out/pred_bands1.log | tail -40
