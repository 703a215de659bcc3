"""file API twice on a c5s-size stack (4096x4096x16, way tiles, auto select) with two noise realisations; prints the engine's error text"""
import ctypes as C, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import lfm_b200 as L
from conftest import lf_synth
L.set_devices(0, 1); L.set_way(0)
base = lf_synth((16, 4096, 4096), 13, seed=12345)
rng = np.random.default_rng(777); m = base.astype(np.float32)
second = np.clip(np.rint(m + rng.normal(0, 1, m.shape).astype(np.float32) * np.sqrt(np.maximum(m, 1)) * 0.5), 0, 65535).astype(np.uint16)
fn = "/dev/shm/repro_c5s.lfm"; out = np.empty_like(base)
xyzct = L._xyzct(base.shape); bs = L._bs((96, 96, 8, 1, 1)); bsp = C.cast(bs, C.c_void_p)
for i, a in enumerate((base, second)):
    rc = L.lib.writeLFMstackEx(a.ctypes.data, os.fsencode(fn), xyzct, 1, -1, None, bsp, 1, None, 0, 13)
    sz = os.path.getsize(fn) if rc == 0 else -1
    dt = C.c_int()
    rc2 = L.lib.readKLBstackInPlace(os.fsencode(fn), out.ctypes.data, C.byref(dt), -1) if rc == 0 else -1
    err = L.lib.lfmLastError(); err = C.cast(err, C.c_char_p).value if err else b""
    st = L.stats()
    if i == 1:                                      # are the streams valid bzip2?  (libbz2 through python)
        import bz2, struct
        raw = open(fn, "rb").read()
        nb = L.lib.lfmNumBlocks(xyzct, bsp)
        offs = np.frombuffer(raw[320:320 + 8 * nb], dtype=np.uint64).astype(np.int64)
        pay = raw[320 + 8 * nb:]
        bad = []
        for b in range(nb):
            beg = int(offs[b - 1]) if b else 0
            try:
                bz2.decompress(pay[beg:int(offs[b])])
            except Exception as ex:
                bad.append((b, beg, int(offs[b]) - beg, repr(ex)[:60]))
        beg = int(offs[nb - 2]); open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "..", "..", "gpurun_out", "last_block_%s.bz2" % os.environ.get("TAG", "x")) if os.environ.get("TAG") else "/tmp/last_block.bz2", "wb").write(pay[beg:int(offs[nb - 1])])
        print("  libbz2 on the %d streams of the file: %d bad %s" % (nb, len(bad), bad[:5]), flush=True)
    print("step %d (predictor %d): write rc %d (file %d bytes), read rc %d, equal %s, error text: %s" % (i, st.predictor, rc, sz, rc2, bool(rc2 == 0 and np.array_equal(out, a)), err), flush=True)
os.remove(fn)
