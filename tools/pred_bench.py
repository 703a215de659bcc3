"""predictor-only timings (lfmDebugPredictDevice): ways x predictors x stack shapes; prints GB/s at 4 B/px and checks the round trip"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lf_synth
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
L.set_devices(0, 1)
shapes = [(32, 2048, 2048, 15), (1, 2048, 2048, 15), (16, 4096, 4096, 13), (64, 1024, 1024, 13)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
if os.environ.get("SHAPES"): shapes = [tuple(int(v) for v in x.split(",")) for x in os.environ["SHAPES"].split(";")]
ways = [int(x) for x in os.environ.get("WAYS", "0,1,2").split(",")]
videos = [int(x) for x in os.environ.get("VIDEOS", "0,1").split(",")]
ks = [int(x) for x in os.environ.get("KS", "4,7").split(",")]
for (F, H, W, T) in shapes:
    base = lf_synth((1, H, W), T)
    big = torch.from_numpy(np.ascontiguousarray(np.tile(base, (F, 1, 1))).view(np.int16)).cuda()
    big += torch.arange(F, dtype=torch.int16, device="cuda").view(F, 1, 1)
    sym = torch.empty_like(big); back = torch.empty_like(big)
    xyz = L._u32x5(W, H, F, 1, 1); ms = C.c_float()
    for way in ways:
        L.set_way(way)
        for k in ks:
            for video in (videos if way == 0 else (0,)):
                res = []
                for inv, src, dst in ((0, big, sym), (1, sym, back)):
                    back.zero_() if inv else None
                    rc = L.lib.lfmDebugPredictDevice(src.data_ptr(), dst.data_ptr(), xyz, T, k, video, inv, 2, C.byref(ms))
                    rc |= L.lib.lfmDebugPredictDevice(src.data_ptr(), dst.data_ptr(), xyz, T, k, video, inv, 5, C.byref(ms))
                    assert rc == 0, rc
                    res.append((ms.value, 4.0 * big.numel() / (ms.value * 1e-3) / 1e9))
                ok = bool(torch.equal(back, big))
                print("%dx%dx%d T=%d way %d k %d video %d: fwd %.3f ms %.0f GB/s | inv %.3f ms %.0f GB/s | round trip %s" % (W, H, F, T, way, k, video, res[0][0], res[0][1], res[1][0], res[1][1], "ok" if ok else "MISMATCH"), flush=True)
    del big, sym, back
