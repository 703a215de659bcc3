"""BASELINE.json configs[2..4] at FULL size on one B200: device-resident compress / decompress throughput, exact round trip,
compression ratio, and (C3, C4) the host-buffer e2e path.  Frames are synthesised on the GPU from one LF-synth frame
(pattern x slow z modulation + Poisson noise), so the 6.7 GB stack of C5 never exists on the host.
  python tools/fullsize.py [c3 c4 c5]      -> one JSON line per config"""
import ctypes as C, importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lf_synth
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
L.set_devices(0, 1)
CFG = {"c3": (101, 2048, 2048, 13, 1, 0, "configs[2]: 2048x2048x101 z-stack, Nnum=13, angle predictor, 2-D entropy selection"),
       "c4": (1000, 1024, 1024, 13, 0, 0x80, "configs[3]: 1024x1024x1000 video stack (as z, SURVEY F.5), way tiles, video bit + selection"),
       "c5": (200, 4096, 4096, 13, 0, 0, "configs[4]: 4096x4096x200 stack, way tiles, selection; full decode")}

def synth(F, H, W, T):
    base = torch.from_numpy(lf_synth((1, H, W), T).astype(np.float32)).cuda()[0]
    out = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(4242)
    for z in range(F):
        lam = base * (1.0 + 0.15 * np.sin(0.05 * z)) + 1.0
        fr = torch.poisson(lam, generator=g).clamp_(0, 65535)
        out[z] = fr.to(torch.int32).to(torch.int16)              # wraps like uint16
    return out

for name in (sys.argv[1:] or ["c3", "c4", "c5"]):
    F, H, W, T, way, hv, desc = CFG[name]
    L.set_way(way)
    d = synth(F, H, W, T); back = torch.empty_like(d)
    raw = d.numel() * 2
    xyzct = L._u32x5(W, H, F, 1, 1); nb = L.lib.lfmNumBlocks(xyzct, None)
    off = np.zeros(nb, np.uint64); shv = C.c_uint8(); dp = C.c_void_p(); pb = C.c_uint64()
    res = {"config": desc, "raw_bytes": raw, "klb_blocks": int(nb)}
    tcs, tds = [], []
    for rep in range(2):                                           # second repetition = warm buffers
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = L.lib.lfmCompressDevice(d.data_ptr(), xyzct, None, hv, T, C.byref(shv), off.ctypes.data, nb, C.byref(dp), C.byref(pb))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        assert rc == 0, (rc, L.lib.lfmLastError())
        sc = L.stats()
        rc = L.lib.lfmDecompressDevice(dp, off.ctypes.data, nb, xyzct, None, shv.value, T, back.data_ptr())
        torch.cuda.synchronize(); t2 = time.perf_counter()
        assert rc == 0, (rc, L.lib.lfmLastError())
        sd = L.stats()
        tcs.append(t1 - t0); tds.append(t2 - t1)
    res.update(round_trip_exact=bool(torch.equal(back, d)), stored_header_version=int(shv.value), compression_ratio=raw / pb.value,
               compress_gbs=raw / min(tcs) / 1e9, decompress_gbs=raw / min(tds) / 1e9,
               compress_stage_ms={k: round(getattr(sc, "ms_" + k), 2) for k in ("select", "predict", "rle", "bwt", "mtf", "huff")},
               decompress_stage_ms={k: round(getattr(sd, "ms_" + k), 2) for k in ("decode", "imtf", "ibwt", "unrle", "unpredict")})
    assert res["round_trip_exact"], name
    if name in ("c3", "c4"):                                       # host buffers through the C ABI (pinned), H2D + D2H inside
        h = d.cpu().pin_memory(); hn = h.numpy().view(np.uint16)
        blob = torch.empty(raw // 2 + raw // 4 + (1 << 20), dtype=torch.uint8).pin_memory(); bn = blob.numpy()
        ho = torch.empty_like(h).pin_memory(); hon = ho.numpy().view(np.uint16)
        best = [1e9, 1e9]
        for rep in range(2):
            t0 = time.perf_counter(); nblob = L.compress_into(hn, bn, header_version=hv, nnum=T, way=way); t1 = time.perf_counter()
            L.decompress_into(bn, nblob, hon, way=way); t2 = time.perf_counter()
            best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
        assert np.array_equal(hon, hn)
        res.update(e2e_compress_gbs=raw / best[0] / 1e9, e2e_decompress_gbs=raw / best[1] / 1e9, file_bytes=int(nblob))
        del h, blob, ho
        # the drop-in file API with PAGEABLE memory (what JNI / MEX callers pass), file on /dev/shm: writeLFMstackEx + readKLBstackInPlace
        pg = d.cpu().numpy().view(np.uint16).copy(); po = np.empty_like(pg)
        fn = os.fsencode("/dev/shm/lfm_fullsize_%d.lfm" % os.getpid())
        bestf = [1e9, 1e9]
        for rep in range(2):
            t0 = time.perf_counter()
            rc = L.lib.writeLFMstackEx(pg.ctypes.data, fn, xyzct, 1, -1, None, None, 1, None, hv, T); t1 = time.perf_counter()
            assert rc == 0, rc
            dt = C.c_int(); rc = L.lib.readKLBstackInPlace(fn, po.ctypes.data, C.byref(dt), -1); t2 = time.perf_counter()
            assert rc == 0, rc
            bestf = [min(bestf[0], t1 - t0), min(bestf[1], t2 - t1)]
        assert np.array_equal(po, pg)
        os.remove(fn)
        res.update(file_api_compress_gbs=raw / bestf[0] / 1e9, file_api_decompress_gbs=raw / bestf[1] / 1e9)
        del pg, po
    print(json.dumps(res), flush=True)
    del d, back
    torch.cuda.empty_cache()
