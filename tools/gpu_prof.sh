#!/bin/bash
# usage: tools/gpu_prof.sh <kernel-regex> <workload> ; plain run first, then one ncu --set full capture of the kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof.py $2 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -f -o gpurun_out/prof_$1 python tools/prof.py $2 2 > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_plain.log; tail -3 gpurun_out/prof_ncu.log; ls -la gpurun_out/*.ncu-rep
