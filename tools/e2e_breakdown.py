"""Where the host-side time of the file API goes at configs[1] (one 2048x2048 frame): wall clock of writeLFMstackEx /
readKLBstackInPlace on a /dev/shm file from pageable numpy memory against the engine's own timers (lfm_stats).
   python tools/e2e_breakdown.py [reps]"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import lfm_b200 as L
from conftest import lf_synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
H = W = 2048; nnum = 15
L.set_way(2)
pool = [lf_synth((1, H, W), nnum, seed=100 + i) for i in range(6)]
out = np.empty_like(pool[0])
fname = "/dev/shm/lfm_e2e_breakdown.lfm"
xyzct = L._xyzct(pool[0].shape); bs = L._bs((96, 96, 1, 1, 1)); bsp = C.cast(bs, C.c_void_p)
fields = ["ms_total", "ms_h2d", "ms_d2h", "ms_select", "ms_predict", "ms_rle", "ms_bwt", "ms_mtf", "ms_huff", "ms_decode", "ms_imtf", "ms_ibwt", "ms_unrle", "ms_unpredict"]
acc = {"w_wall": 0.0, "r_wall": 0.0}
for f in fields: acc["w_" + f] = 0.0; acc["r_" + f] = 0.0
for i in range(reps + 3):
    a = pool[i % len(pool)]
    t0 = time.perf_counter()
    rc = L.lib.writeLFMstackEx(a.ctypes.data, os.fsencode(fname), xyzct, 1, -1, None, bsp, 1, None, 12, nnum); assert rc == 0
    t1 = time.perf_counter()
    sw = L.stats()
    dt = C.c_int()
    t2 = time.perf_counter()
    rc = L.lib.readKLBstackInPlace(os.fsencode(fname), out.ctypes.data, C.byref(dt), -1); assert rc == 0
    t3 = time.perf_counter()
    sr = L.stats()
    if i >= 3:
        acc["w_wall"] += (t1 - t0) * 1e3; acc["r_wall"] += (t3 - t2) * 1e3
        for f in fields: acc["w_" + f] += getattr(sw, f); acc["r_" + f] += getattr(sr, f)
assert np.array_equal(out, pool[(reps + 2) % len(pool)])
os.remove(fname)
for side in "wr":
    k = sum(acc[side + "_" + f] for f in fields[3:]) / reps
    print("%s: wall %.3f ms | engine total %.3f | h2d %.3f | kernels %.3f | d2h %.3f | outside the engine timers %.3f" % (
        "write" if side == "w" else "read ", acc[side + "_wall"] / reps, acc[side + "_ms_total"] / reps, acc[side + "_ms_h2d"] / reps, k,
        acc[side + "_ms_d2h"] / reps, (acc[side + "_wall"] - acc[side + "_ms_total"]) / reps))
