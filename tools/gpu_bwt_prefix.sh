#!/bin/bash
# usage: tools/gpu_bwt_prefix.sh: stage times of c2 / c3s for radix prefixes of 8, 6, 5, 4 bytes (LFM_B200_BWT_PREFIX), parity tests at 6
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for p in 8 6 5 4; do
  LFM_B200_BWT_PREFIX=$p python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p$p.json 2>/dev/null
  echo "c2 prefix $p: $(python -c "import json;d=json.loads(open('gpurun_out/bench_p$p.json').read().strip().splitlines()[-1]);print('value',round(d['value'],3),'bwt',round(d['stage_ms_per_step']['bwt'],3))")"
done
for p in 8 6 4; do
  LFM_B200_BWT_PREFIX=$p python bench.py --steps 3 --warmup 2 --workload c3s > gpurun_out/bench_c3s_p$p.json 2>/dev/null
  echo "c3s prefix $p: $(python -c "import json;d=json.loads(open('gpurun_out/bench_c3s_p$p.json').read().strip().splitlines()[-1]);print('value',round(d['value'],3),'bwt',round(d['stage_ms_per_step']['bwt'],3))")"
done
LFM_B200_BWT_PREFIX=6 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
