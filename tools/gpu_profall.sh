#!/bin/bash
# usage: tools/gpu_profall.sh <workload> [skip] [count]; plain run first, then ONE ncu --set full capture of the kernels of the
# 2nd round trip, summarised on the box into gpurun_out/profall_<wl>.md (the report itself is too big to bring back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
wl=${1:-c2}; skip=${2:-14}; cnt=${3:-14}
python tools/prof.py $wl 2 > gpurun_out/profall_${wl}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_' -s $skip -c $cnt -f -o /tmp/profall_$wl python tools/prof.py $wl 2 > gpurun_out/profall_${wl}_ncu.log 2>&1
tail -2 gpurun_out/profall_${wl}_plain.log; tail -2 gpurun_out/profall_${wl}_ncu.log
python tools/ncu_report.py /tmp/profall_$wl.ncu-rep gpurun_out/profall_$wl.md 24
