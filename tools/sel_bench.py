"""mode selection alone (lfmSelectDevice) on one 2048x2048 frame, way space: python tools/sel_bench.py [reps]"""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lf_synth
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
L.set_devices(0, 1); L.set_way(2)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
d = torch.from_numpy(lf_synth((1, 2048, 2048), 15)[0].view(np.int16)).cuda()
k = C.c_int(); e = (C.c_float * 8)()
for i in range(reps + 2):
    if i == 2: torch.cuda.synchronize(); t0 = time.perf_counter()
    assert L.lib.lfmSelectDevice(d.data_ptr(), (C.c_uint32 * 2)(2048, 2048), 15, C.byref(k), e) == 0
torch.cuda.synchronize()
print("select: %.3f ms per call, winner %d, entropies %s" % ((time.perf_counter() - t0) / reps * 1e3, k.value, [round(x, 4) for x in e]))
