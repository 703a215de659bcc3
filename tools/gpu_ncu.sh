#!/bin/bash
# launch list (per-kernel device time) of the default bench command; plain run first, ncu only if it exits 0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log | cut -c1-300
wc -l gpurun_out/launches.csv
