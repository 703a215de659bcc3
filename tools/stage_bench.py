"""stage times of one device-resident round trip of a workload (c2 / c3s / c5s), a few repetitions: quick A/B of kernel switches"""
import ctypes as C, importlib, sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lf_synth
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
wl = sys.argv[1] if len(sys.argv) > 1 else "c3s"
nfr, H, W, nnum, way, hv = {"c2": (1, 2048, 2048, 15, 2, 12), "c3s": (16, 2048, 2048, 13, 1, 0), "c5s": (16, 4096, 4096, 13, 0, 0)}[wl]
L.set_devices(0, 1); L.set_way(way)
a = lf_synth((nfr, H, W), nnum)
d = torch.from_numpy(a.view(np.int16)).cuda(); out = torch.empty_like(d)
xyzct = L._u32x5(W, H, nfr, 1, 1); nb = L.lib.lfmNumBlocks(xyzct, None)
off = np.zeros(nb, np.uint64); shv = C.c_uint8(); dp = C.c_void_p(); pb = C.c_uint64()
acc = {}
for i in range(5):
    assert L.lib.lfmCompressDevice(d.data_ptr(), xyzct, None, hv, nnum, C.byref(shv), off.ctypes.data, nb, C.byref(dp), C.byref(pb)) == 0
    sc = L.stats()
    assert L.lib.lfmDecompressDevice(dp, off.ctypes.data, nb, xyzct, None, shv.value, nnum, out.data_ptr()) == 0
    sd = L.stats()
    if i >= 2:
        for k in ("rle", "bwt", "mtf", "huff"): acc[k] = acc.get(k, 0) + getattr(sc, "ms_" + k) / 3
        for k in ("decode", "imtf", "ibwt", "unrle", "unpredict"): acc[k] = acc.get(k, 0) + getattr(sd, "ms_" + k) / 3
torch.cuda.synchronize(); assert torch.equal(out, d)
print(wl, {k: round(v, 3) for k, v in acc.items()})
print("select ms", round(sc.ms_select, 3), "selected", sc.selected, "predictor", sc.predictor, [round(x, 4) for x in sc.entropy])
