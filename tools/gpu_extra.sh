#!/bin/bash
# extra coverage: the optional code paths through the gpu test-suite, then the full-size configs
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for envs in "LFM_B200_GROUPS=3" "LFM_B200_BANDS=2 LFM_B200_BANDS_MIN=1" "LFM_B200_BANDS=0"; do
  echo "== $envs"; env $envs timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -3
done > gpurun_out/pytest_variants.log 2>&1
cat gpurun_out/pytest_variants.log
timeout 900 python tools/fullsize.py c3 c4 c5 > gpurun_out/fullsize.jsonl 2> gpurun_out/fullsize.err; cat gpurun_out/fullsize.jsonl | cut -c1-700; tail -3 gpurun_out/fullsize.err
