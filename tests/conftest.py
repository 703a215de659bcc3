"""Shared test plumbing: markers, the oracle loader (test infrastructure only), synthetic light-field data."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def build_oracle():
    """compile oracle/*.c into oracle/liblfm_oracle.so (gcc, a second or two)"""
    so = os.path.join(ORACLE_DIR, "liblfm_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("lfm_oracle.c", "bz2_oracle.c")]
    deps = srcs + [os.path.join(ORACLE_DIR, "bz2_randtable.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in deps):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so] + srcs + ["-lm"])
    return so


class Oracle:
    """ctypes view of oracle/liblfm_oracle.so"""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.bz2o_compress.restype = C.c_size_t
        L.bz2o_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.bz2o_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.bz2o_trace_new.restype = C.c_void_p
        L.bz2o_trace_delete.argtypes = [C.c_void_p]
        L.bz2o_trace_nblocks.argtypes = [C.c_void_p]
        L.bz2o_trace_block.restype = C.c_void_p; L.bz2o_trace_block.argtypes = [C.c_void_p, C.c_int]
        L.bz2o_blk_i32.argtypes = [C.c_void_p, C.c_int]; L.bz2o_blk_i32.restype = C.c_int32
        L.bz2o_blk_ptr.argtypes = [C.c_void_p, C.c_int]; L.bz2o_blk_ptr.restype = C.c_void_p
        L.lfmo_entropy2d.restype = C.c_float
        L.lfmo_entropy2d.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.lfmo_predict_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6
        L.lfmo_unpredict_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6
        L.lfmo_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.lfmo_write.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        L.lfmo_read.argtypes = [C.c_char_p, C.c_void_p, C.c_int]

    def bz2_compress(self, data, level, trace=False):
        data = bytes(data)
        cap = len(data) * 2 + 1000
        buf = C.create_string_buffer(cap)
        tr = self.lib.bz2o_trace_new() if trace else None
        n = self.lib.bz2o_compress(data, len(data), level, buf, cap, tr)
        assert n <= cap
        out = buf.raw[:n]
        if not trace:
            return out
        blocks = []
        for i in range(self.lib.bz2o_trace_nblocks(tr)):
            b = self.lib.bz2o_trace_block(tr, i)
            g = lambda w: self.lib.bz2o_blk_i32(b, w)
            nblock, n_mtf, n_sel = g(0), g(4), g(6)
            arr = lambda w, cnt, ty: np.ctypeslib.as_array(C.cast(self.lib.bz2o_blk_ptr(b, w), C.POINTER(ty)), (cnt,)).copy()
            blocks.append(dict(nblock=nblock, crc=g(1) & 0xFFFFFFFF, orig_ptr=g(2), n_in_use=g(3), n_mtf=n_mtf, n_groups=g(5),
                               n_sel=n_sel, periodic=g(7), rle1=arr(0, nblock, C.c_uint8), bwt=arr(1, nblock, C.c_uint8),
                               mtfv=arr(2, n_mtf, C.c_uint16), selector=arr(3, n_sel, C.c_uint8)))
        self.lib.bz2o_trace_delete(tr)
        return out, blocks

    def set_randomised(self, on):
        """test hook: blocks are written the way bzip2 <= 0.9.0 did after a failed sort (randomised bit set, bytes XORed with the
        BZ2_rNums mask before the sort); libbz2 decodes such streams, nothing written since 0.9.5 produces them"""
        self.lib.bz2o_set_randomised(1 if on else 0)

    def bz2_decompress(self, data, cap):
        buf = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        rc = self.lib.bz2o_decompress(bytes(data), len(data), buf, cap, C.byref(n))
        return rc, buf.raw[:n.value]

    def predict_frame(self, cur, prev, T, way, k, zflag):
        cur = np.ascontiguousarray(cur, np.uint16); H, W = cur.shape
        prev_p = np.ascontiguousarray(prev, np.uint16).ctypes.data if prev is not None else None
        sym = np.empty((H, W), np.uint16)
        rc = self.lib.lfmo_predict_frame(cur.ctypes.data, prev_p, sym.ctypes.data, W, H, T, way, k, zflag)
        assert rc == 0
        return sym

    def select(self, frame, T, way):
        frame = np.ascontiguousarray(frame, np.uint16); H, W = frame.shape
        e = (C.c_float * 8)()
        k = self.lib.lfmo_select(frame.ctypes.data, W, H, T, way, e)
        return k, list(e)

    def write(self, img, filename, hv, nnum, way, block_size=None):
        img = np.ascontiguousarray(img, np.uint16)
        s = list(img.shape)
        while len(s) < 5:
            s.insert(0, 1)
        xyzct = (C.c_uint32 * 5)(s[4], s[3], s[2], s[1], s[0])
        bs = (C.c_uint32 * 5)(*block_size) if block_size is not None else None
        shv = C.c_int(-1)
        rc = self.lib.lfmo_write(img.ctypes.data, os.fsencode(filename), xyzct, bs, hv, nnum, way, C.byref(shv), None)
        return rc, shv.value

    def read(self, filename, shape, way):
        out = np.empty(shape, np.uint16)
        rc = self.lib.lfmo_read(os.fsencode(filename), out.ctypes.data, way)
        return rc, out


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


def lf_synth(shape_zyx, nnum, seed=12345):
    """LF-synth v1: Poisson-like light field with microlens vignetting, deterministic."""
    Z, Y, X = shape_zyx
    rng = np.random.default_rng(seed)
    x = np.arange(X)[None, None, :].astype(np.float64); y = np.arange(Y)[None, :, None].astype(np.float64)
    z = np.arange(Z)[:, None, None].astype(np.float64)
    S = 400 + 300 * np.sin(x / 97 + 0.05 * z) * np.cos(y / 131) + 200 * np.sin((x + y) / 37)
    u = (np.arange(X) % nnum - (nnum - 1) / 2)[None, None, :]; v = (np.arange(Y) % nnum - (nnum - 1) / 2)[None, :, None]
    V = np.exp(-(u * u + v * v) / (0.18 * nnum * nnum))
    m = 100 + S * V
    img = np.clip(np.rint(rng.normal(m, np.sqrt(m))), 0, 65535).astype(np.uint16)
    return img


def lf_synth_int(shape_zyx, nnum, seed=1, device="cpu", z0=0):
    """Full-size deterministic light-field stacks (torch tensor, int16 view of uint16 pixels): the LF-synth pattern in INTEGER
    arithmetic only (triangle waves, quadratic microlens vignetting, hashed noise scaled by an integer square root), so that the
    GPU box generates in a second exactly the bytes the committed md5s of tests/golden/fullsize.json were computed on.
    Frames z0 .. z0 + Z - 1 of the (unbounded) stack."""
    import torch
    Z, Y, X = shape_zyx
    out = torch.empty((Z, Y, X), dtype=torch.int16, device=device)
    x = torch.arange(X, dtype=torch.int64, device=device)[None, None, :]
    y = torch.arange(Y, dtype=torch.int64, device=device)[None, :, None]
    tri = lambda v, p: ((v % p) - p // 2).abs()
    u2 = 2 * (x % nnum) - (nnum - 1); v2 = 2 * (y % nnum) - (nnum - 1)
    V = (256 - (u2 * u2 * 160) // (nnum * nnum)) * (256 - (v2 * v2 * 160) // (nnum * nnum))          # <= 65536
    M31 = (1 << 31) - 1
    step = max(1, (1 << 24) // (X * Y))
    for a in range(0, Z, step):
        b = min(Z, a + step)
        z = torch.arange(z0 + a, z0 + b, dtype=torch.int64, device=device)[:, None, None]
        S = 250 + 3 * tri(x + 5 * z, 194) + 2 * tri(y + z, 262) + tri(x + y, 74) * 4
        m = 100 + (S * V >> 16)
        g = torch.zeros((b - a, Y, X), dtype=torch.int64, device=device)
        h = (x * 73856093 + y * 19349663 + z * 83492791 + seed * 2654435761) & M31
        for _ in range(4):
            h = (h * 1103515245 + 12345) & M31
            h = h ^ (h >> 13)
            g += (h >> 7) & 255
        r = torch.sqrt(m.to(torch.float64)).to(torch.int64)                                           # exact: m < 2^52
        pix = (m + ((g - 510) * r) // 148).clamp_(0, 65535)
        out[a:b] = (pix - ((pix >> 15) << 16)).to(torch.int16)                                        # uint16 bits in an int16 tensor
    return out


def load_reference_gpu(way):
    """the UNMODIFIED reference built for the GPU (oracle/_ref/liblfmref_gpu_way<w>.so, oracle/build_ref.py): its CUDA predictor,
    thrust sort / reduce and threaded CPU bzip2 exactly as shipped.  None when the library (or a CUDA device) is absent."""
    so = os.path.join(ORACLE_DIR, "_ref", "liblfmref_gpu_way%d.so" % way)
    if not os.path.exists(so) or not has_cuda():
        return None
    try:
        lib = C.CDLL(so)
    except OSError:
        return None
    lib.ref_write.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.ref_read_full.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    return lib


def golden_img_tif():
    """the reference's only fixture, testData/img.tif (101x151x29 uint16), stored as npz (tests/golden/make_golden.py)"""
    return np.load(os.path.join(GOLDEN, "img_tif_101x151x29_u16.npz"))["img"]


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def multiblock_cases():
    """byte strings (even length) whose bzip2 stream at the KLB level rule -- level = min(9, ceil(len / 100000)), one bzip2 block
    holds 100000 * level - 19 run-length coded bytes -- has more than one block, or sits exactly on the rules that decide it
    (bzlib.c:216-340: a block is closed by the record that fills it, unless only the last input byte is left)"""
    rng = np.random.default_rng(77)
    def norun(n):                     # no two equal neighbours: post-RLE1 length == n
        a = rng.integers(0, 255, n, dtype=np.uint8)
        a[1:][a[1:] == a[:-1]] += 1
        for i in np.nonzero(a[1:] == a[:-1])[0]:
            a[i + 1] = (int(a[i]) + 7) % 256
        return a
    M2 = 199981                       # level 2
    c = {}
    c["l2_exact_plus1"] = norun(M2 + 1).tobytes()                    # one byte past a full block: stays ONE block
    c["l2_exact_plus3"] = norun(M2 + 3).tobytes()                    # second block of 3 bytes
    t = norun(M2 + 3); t[-1] = t[-2]; c["l2_tail_run2"] = t.tobytes()   # what is left after the boundary is one 2-byte record
    t = norun(M2 + 5); t[M2 - 3:M2 + 1] = 9; t[M2 - 4] = 8; t[M2 + 1] = 10
    c["l2_run4_across"] = t.tobytes()                                 # a 4-run (5 output bytes) straddles the block limit
    c["l2_runs4"] = np.repeat(rng.integers(0, 256, 50000, dtype=np.uint8), 4).tobytes()      # 200000 -> 250000 bytes: two blocks
    c["l2_runs300"] = np.repeat(rng.integers(0, 256, 666, dtype=np.uint8), 300)[:199800].tobytes()  # long runs, 255-records
    c["l1_poisson"] = rng.poisson(4, 49990).astype(np.uint16).tobytes()                       # level 1, just below one block
    t = rng.integers(0, 256, 99996, dtype=np.uint8); c["l1_random_2blocks"] = t.tobytes()   # level 1: 99996 > 99981
    c["l9_2MB_3blocks"] = rng.poisson(30, 1000000).astype(np.uint16).tobytes()              # level 9, three blocks
    c["l9_1MB_runs"] = np.repeat(rng.integers(0, 256, 250000, dtype=np.uint8), 4).tobytes()  # level 9: 1 MB -> 1.25 MB, two blocks
    return c
