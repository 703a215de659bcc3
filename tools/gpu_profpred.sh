#!/bin/bash
# usage: tools/gpu_profpred.sh <kernel-regex> [skip] ; ncu --set full of one predictor kernel launch of tools_pred_bench.py (env SHAPES/KS/WAYS/VIDEOS), summarised on the box
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
kre=${1:-k_unpredict_bands}; skip=${2:-1}
tag=$(echo "$kre" | tr -c 'a-zA-Z0-9_\n' '_')
python tools/pred_bench.py > gpurun_out/profpred_${tag}_plain.log 2>&1 && tail -3 gpurun_out/profpred_${tag}_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c 1 -f -o /tmp/profpred_$tag python tools/pred_bench.py > gpurun_out/profpred_${tag}.log 2>&1
tail -2 gpurun_out/profpred_${tag}.log
python tools/ncu_report.py /tmp/profpred_$tag.ncu-rep gpurun_out/profpred_$tag.md 40
