"""configs[4] decode path with klb_ROI crops (readImage, src/klb_imageIO.cpp:2614-2682) on a 4096x4096x16 slice written to a file:
full read, one XY plane, one 512x512x16 box -- with the selected predictor (whole frames of the touched slabs are decoded, then
cropped) and with the predictor off (only the KLB blocks that intersect the ROI are read from the file and decoded)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
from conftest import lf_synth
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
L.set_devices(0, 1)
F, H, W, T = 16, 4096, 4096, 13
rng = np.random.default_rng(3)
base = lf_synth((1, H, W), T)[0].astype(np.float32)
a = np.empty((F, H, W), np.uint16)
for z in range(F):
    a[z] = np.clip(np.rint(base * (1 + 0.1 * np.sin(0.3 * z)) + rng.normal(0, 1, (H, W)).astype(np.float32) * np.sqrt(base) * 0.7), 0, 65535).astype(np.uint16)
tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
out = []
for hv, label in ((0, "auto-selected predictor, way tiles"), (8, "predictor off")):
    fn = os.path.join(tmpdir, "lfm_roi_%d_%d.lfm" % (os.getpid(), hv))
    t0 = time.perf_counter(); L.write_stack(a, fn, header_version=hv, nnum=T, way=0); tw = time.perf_counter() - t0
    fsz = os.path.getsize(fn)
    res = {"file": label, "file_bytes": fsz, "write_file_gbs": a.nbytes / tw / 1e9}
    L.read_stack(fn, way=0)                                              # warm buffers
    for name, lb, ub in (("full", (0, 0, 0, 0, 0), (W - 1, H - 1, F - 1, 0, 0)), ("xy_plane_z7", (0, 0, 7, 0, 0), (W - 1, H - 1, 7, 0, 0)),
                         ("box_512x512x16", (1000, 2000, 0, 0, 0), (1511, 2511, F - 1, 0, 0)), ("box_512x512x1", (1000, 2000, 9, 0, 0), (1511, 2511, 9, 0, 0))):
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter(); r = L.read_roi(fn, lb, ub, way=0); best = min(best, time.perf_counter() - t0)
        want = a[lb[2]:ub[2] + 1, lb[1]:ub[1] + 1, lb[0]:ub[0] + 1]
        assert np.array_equal(r[0, 0], want), (label, name)
        st = L.stats()
        res[name] = {"ms": round(best * 1e3, 2), "roi_mbytes": round(want.nbytes / 1e6, 2), "roi_gbs": round(want.nbytes / best / 1e9, 3),
                     "kernel_launches": int(st.gpu_launches)}
    os.remove(fn)
    out.append(res)
    print(json.dumps(res), flush=True)
