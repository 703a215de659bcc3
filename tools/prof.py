"""small driver for ncu captures: a few device-resident round trips of one workload"""
import ctypes as C, importlib, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lf_synth
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nfr, H, W, nnum, way, hv = {"c2": (1, 2048, 2048, 15, 2, 12), "c3s": (16, 2048, 2048, 13, 1, 0), "c3f": (16, 2048, 2048, 13, 1, 12), "c4s": (32, 1024, 1024, 13, 0, 0x80 | 12)}[wl]
L.set_devices(0, 1); L.set_way(way)
a = lf_synth((nfr, H, W), nnum)
d = torch.from_numpy(a.view(np.int16)).cuda(); out = torch.empty_like(d)
xyzct = L._u32x5(W, H, nfr, 1, 1); nb = L.lib.lfmNumBlocks(xyzct, None)
off = np.zeros(nb, np.uint64); shv = C.c_uint8(); dp = C.c_void_p(); pb = C.c_uint64()
for i in range(reps):
    assert L.lib.lfmCompressDevice(d.data_ptr(), xyzct, None, hv, nnum, C.byref(shv), off.ctypes.data, nb, C.byref(dp), C.byref(pb)) == 0
    assert L.lib.lfmDecompressDevice(dp, off.ctypes.data, nb, xyzct, None, shv.value, nnum, out.data_ptr()) == 0
torch.cuda.synchronize()
assert torch.equal(out, d)
print("ok", pb.value)
