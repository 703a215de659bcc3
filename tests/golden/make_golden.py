#!/usr/bin/env python3
"""Generate tests/golden/* from the REAL reference (run in the build container, where /root/reference exists).

  img_tif_101x151x29_u16.npz   pixel data of the reference's only fixture testData/img.tif (PIL decode)
  sample{1,2,3}.bz2            the reference's bzip2 known-answer files (src/external/bzip2-1.0.6/sample*.bz2);
                               the .ref inputs are recovered by decompressing them, levels are 1/2/3 (Makefile:58-69)
  golden.json                  md5 / size / stored headerVersion of .lfm files written by the UNMODIFIED reference
                               (oracle/_ref/liblfmref_cpu_way{0,1,2}.so = its own sources, CUDA kernels run on the CPU),
                               entropies of the 8 candidates, per way, for img.tif and small LF-synth stacks.
"""
import ctypes as C, hashlib, json, os, shutil, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lf_synth  # noqa: E402

REF = "/root/reference"

def main():
    from PIL import Image
    im = Image.open(os.path.join(REF, "testData/img.tif")); fr = []
    for i in range(im.n_frames):
        im.seek(i); fr.append(np.array(im))
    img = np.stack(fr).astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "img_tif_101x151x29_u16.npz"), img=img)
    for i in (1, 2, 3):
        shutil.copy(os.path.join(REF, "src/external/bzip2-1.0.6/sample%d.bz2" % i), os.path.join(HERE, "sample%d.bz2" % i))
        os.chmod(os.path.join(HERE, "sample%d.bz2" % i), 0o644)
    refs = [C.CDLL(os.path.join(ROOT, "oracle/_ref/liblfmref_cpu_way%d.so" % w)) for w in range(3)]
    for r in refs:
        r.ref_entropy_2d.restype = C.c_float
    out = {"files": [], "entropy": []}
    stacks = {"img_tif": (img, 13), "synth_40x70x90_n15": (lf_synth((40, 70, 90), 15), 15), "synth_1x200x230_n13": (lf_synth((1, 200, 230), 13), 13),
              "synth_9x64x64_n11": (lf_synth((9, 64, 64), 11), 11)}
    tmp = "/tmp/golden_ref.lfm"
    for name, (a, nnum) in stacks.items():
        Z, H, W = a.shape
        xyzct = (C.c_uint32 * 5)(W, H, Z, 1, 1)
        for way in range(3):
            hvs = [8, 0] + list(range(9, 16)) + ([0x80] + [0x88 + k for k in (1, 2, 3, 5, 6, 7)] if way == 0 and Z > 1 else [])
            for hv in hvs:
                shv = C.c_int()
                err = refs[way].ref_write(a.ctypes.data_as(C.c_void_p), tmp.encode(), xyzct, None, hv, nnum, -1, C.byref(shv))
                data = open(tmp, "rb").read()
                out["files"].append(dict(stack=name, way=way, hv_in=hv, nnum=nnum, err=err, hv_stored=shv.value, size=len(data),
                                         md5=hashlib.md5(data).hexdigest()))
        # candidate entropies of frame 0 via the reference's bwt_entropy_2D on the reference's own symbol images
        f0 = np.ascontiguousarray(a[min(10, Z - 1)]); both = np.ascontiguousarray(np.stack([f0, f0]))
        for way in range(3):
            es = []
            for k in range(8):
                if k == 0:
                    s = f0.copy()
                else:
                    s = np.zeros((H, W), np.uint16)
                    refs[0].ref_predict_frame(both.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), W, H, nnum, way, k, 0)
                es.append(float(refs[way].ref_entropy_2d(s.ctypes.data_as(C.c_void_p), C.c_uint64(W * H), k)))
            out["entropy"].append(dict(stack=name, frame=min(10, Z - 1), way=way, nnum=nnum, e=es))
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", len(out["files"]), "file goldens,", len(out["entropy"]), "entropy goldens")

if __name__ == "__main__":
    main()
