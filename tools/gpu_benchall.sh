#!/bin/bash
# bench lines of every workload (c2 default, c3s, c4s, c5s) + reference arm
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err
for wl in c3s c4s c5s; do python bench.py --steps 4 --warmup 3 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -2 gpurun_out/bench_$wl.err; done
python - <<'PY'
import json
for wl in ("c2","c3s","c4s","c5s"):
    try:
        d=json.loads(open("gpurun_out/bench_%s.json"%wl).read().strip().splitlines()[-1])
        print(wl, "value %.3f comp %.3f decomp %.3f e2e %.3f ratio %.3f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["compression_ratio"]), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "cpu", d["cpu_baseline"]["value"])
    except Exception as e: print(wl, "ERR", e)
PY
