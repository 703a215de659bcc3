import json, os, sys
d0 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
for f in sys.argv[1:] or ("bench_c2", "bench_c2_n2"):
    try:
        d = json.loads(open(os.path.join(d0, f + ".json")).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, {k: (round(d[k], 3) if isinstance(d.get(k), float) else d.get(k)) for k in ("value", "ms_per_step", "compress_gbs", "decompress_gbs", "n_gpus", "gpu_launches")})
    e = d.get("e2e") or {}
    print("  e2e", {k: round(e[k], 3) for k in ("value", "compress_gbs", "decompress_gbs") if k in e}, "mem", {k: round(v, 3) for k, v in (d.get("e2e_memory") or {}).items() if isinstance(v, float)})
    print("  stage", {k: round(v, 3) for k, v in (d.get("stage_ms_per_step") or {}).items()})
    r = d.get("roofline") or {}
    print("  roofline", r.get("kernel", "")[:40], r.get("frac"), " pred", {k: round(v, 3) for k, v in ((d.get("predictor_roofline") or {}).get("frac") or {}).items()})
    c = d.get("cpu_baseline")
    if c: print("  cpu", c.get("value"), c.get("cores"))
