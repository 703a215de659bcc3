#!/bin/bash
# usage: tools/gpu_profk.sh <workload> <kernel-regex> [skip] [count]; ncu --set full of the matching kernels, summarised on the box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
wl=${1:-c2}; kre=${2:-k_}; skip=${3:-1}; cnt=${4:-1}
tag=$(echo "$kre" | tr -c 'a-zA-Z0-9_\n' '_')
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -f -o /tmp/profk_${wl}_$tag python tools/prof.py $wl 2 > gpurun_out/profk_${wl}_${tag}.log 2>&1
tail -2 gpurun_out/profk_${wl}_${tag}.log
python tools/ncu_report.py /tmp/profk_${wl}_$tag.ncu-rep gpurun_out/profk_${wl}_$tag.md ${5:-30}
