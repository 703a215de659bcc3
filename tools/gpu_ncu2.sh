#!/bin/bash
# launch list of the small profiling driver for a workload: tools/gpu_ncu2.sh <workload>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof.py $1 3 > gpurun_out/ncu2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv python tools/prof.py $1 3 > gpurun_out/ncu2_run.log 2>&1
tail -2 gpurun_out/ncu2_run.log
