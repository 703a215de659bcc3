"""Pin oracle/lfm_oracle.c: against committed outputs of the REAL reference (tests/golden/golden.json) and, when
oracle/_ref is built, against the reference library itself (CPU only)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_img_tif, lf_synth

G = json.load(open(os.path.join(GOLDEN, "golden.json")))
STACKS = {"img_tif": lambda: golden_img_tif(), "synth_40x70x90_n15": lambda: lf_synth((40, 70, 90), 15),
          "synth_1x200x230_n13": lambda: lf_synth((1, 200, 230), 13), "synth_9x64x64_n11": lambda: lf_synth((9, 64, 64), 11)}
_cache = {}


def stack(name):
    if name not in _cache:
        _cache[name] = STACKS[name]()
    return _cache[name]


def test_golden_covers_config1():
    """config 1 (img.tif, predictor off) -> 382142 bytes, md5 b8ebe10e..."""
    e = [f for f in G["files"] if f["stack"] == "img_tif" and f["hv_in"] == 8 and f["way"] == 0][0]
    assert e["size"] == 382142 and e["md5"] == "b8ebe10e4beac6bd2b4f583fb0f1e933" and e["hv_stored"] == 0


@pytest.mark.parametrize("idx", range(len(G["files"])))
def test_oracle_file_equals_reference_file(oracle, tmp_path, idx):
    e = G["files"][idx]
    a = stack(e["stack"])
    fn = str(tmp_path / "o.lfm")
    rc, shv = oracle.write(a, fn, e["hv_in"], e["nnum"], e["way"])
    assert rc == e["err"] == 0 and shv == e["hv_stored"]
    data = open(fn, "rb").read()
    assert len(data) == e["size"] and hashlib.md5(data).hexdigest() == e["md5"]
    if idx % 4 == 0:      # the oracle's true inverse restores the input
        rc, back = oracle.read(fn, a.shape, e["way"])
        assert rc == 0 and np.array_equal(back, a)


@pytest.mark.parametrize("idx", range(len(G["entropy"])))
def test_oracle_entropy_equals_reference(oracle, idx):
    e = G["entropy"][idx]
    f0 = np.ascontiguousarray(stack(e["stack"])[e["frame"]])
    for k in range(8):
        s = f0 if k == 0 else oracle.predict_frame(f0, None, e["nnum"], e["way"], k, 0)
        got = oracle.lib.lfmo_entropy2d(np.ascontiguousarray(s).ctypes.data, s.size, k)
        assert abs(got - e["e"][k]) <= 2e-6 * max(1.0, abs(e["e"][k])), (k, got, e["e"][k])


def test_oracle_predictors_against_reference_library(oracle):
    so = os.path.join(ROOT, "oracle", "_ref", "liblfmref_cpu_way0.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    ref = C.CDLL(so)
    rng = np.random.default_rng(7)
    for (W, H, T) in [(47, 38, 5), (30, 17, 13), (13, 13, 13), (14, 27, 13), (64, 48, 15)]:
        for img in (rng.integers(0, 3000, (2, H, W)).astype(np.uint16), rng.integers(0, 65536, (2, H, W)).astype(np.uint16)):
            for way in range(3):
                for k in range(1, 8):
                    for zf in ((0, 1) if way == 0 else (0,)):
                        want = np.zeros((H, W), np.uint16)
                        ref.ref_predict_frame(img.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), W, H, T, way, k, zf)
                        got = oracle.predict_frame(img[0], img[1], T, way, k, zf)
                        assert np.array_equal(got, want), (W, H, T, way, k, zf)
                        back = np.empty((H, W), np.uint16)
                        oracle.lib.lfmo_unpredict_frame(got.ctypes.data, img[1].ctypes.data, back.ctypes.data, W, H, T, way, k, zf)
                        assert np.array_equal(back, img[0])


def test_selection_tie_rule(oracle):
    """all-zero frame: every entropy is 0, the std::map overwrite makes the LARGEST id win (klb_imageIO.cpp:2300-2305)"""
    k, e = oracle.select(np.zeros((40, 52), np.uint16), 13, 0)
    assert k == 7 and all(x == 0 for x in e)
