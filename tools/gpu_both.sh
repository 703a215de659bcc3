#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3f comp %.3f decomp %.3f e2e %.3f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"]), d["stage_ms_per_step"], "cpu", d["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
if [ "$1" == "c3s" ]; then python bench.py --steps 4 --warmup 3 --workload c3s > gpurun_out/bench_c3s.json 2> gpurun_out/bench_c3s.err; python - <<'PY'
import json
f = "gpurun_out/bench_c3s.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.3f comp %.3f decomp %.3f e2e %.3f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"]), d["stage_ms_per_step"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
fi
