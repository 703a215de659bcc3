import importlib, sys, os, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import lf_synth, golden_img_tif
L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
L.set_devices(0, 1)
stacks = [lf_synth((1, 200, 230), 13), golden_img_tif()[10:11], lf_synth((40, 70, 90), 15)[10:11]]
for si, a in enumerate(stacks):
    for way in range(3):
        seen = {}
        for it in range(40):
            L.compress_to_bytes(a, header_version=0, nnum=13, way=way)
            e = tuple(round(float(x), 6) for x in L.stats().entropy)
            seen[e] = seen.get(e, 0) + 1
            # interleave a different-size job to shake buffers
            if it % 3 == 0:
                L.compress_to_bytes(stacks[(si + 1) % 3], header_version=0, nnum=13, way=way)
        print(si, way, len(seen), list(seen.items())[:3])
