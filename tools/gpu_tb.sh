#!/bin/bash
# tests, then bench of the workloads named on the command line
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for wl in "$@"; do st=4; [ $wl == c2 ] && st=10; python bench.py --steps $st --warmup 3 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -2 gpurun_out/bench_$wl.err; done
python - "$@" <<'PY'
import json, sys
for wl in sys.argv[1:]:
    try:
        d=json.loads(open("gpurun_out/bench_%s.json"%wl).read().strip().splitlines()[-1])
        print(wl, "value %.3f comp %.3f decomp %.3f e2e %.3f ratio %.3f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["compression_ratio"]), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
    except Exception as e: print(wl, "ERR", e)
PY
