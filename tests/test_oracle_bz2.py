"""Pin oracle/bz2_oracle.c against the reference's own known-answer vectors and an independent libbz2 (CPU only)."""
import bz2
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _cases():
    rng = np.random.default_rng(1)
    return {
        "empty": b"", "one": b"a", "rand30k": rng.integers(0, 256, 30000, dtype=np.uint8).tobytes(),
        "zeros18432": bytes(18432), "zeros1020": bytes(1020), "const300": np.full(73728, 300, np.uint16).tobytes(),
        "const65535": np.full(9216, 65535, np.uint16).tobytes(),
        "poisson147k": rng.poisson(20, 73728).astype(np.uint16).tobytes(),
        "poisson18k": rng.poisson(3, 9216).astype(np.uint16).tobytes(),
        "two_blocks_100k": rng.integers(0, 4, 100000, dtype=np.uint8).tobytes(),
        "runs4": np.repeat(rng.integers(0, 256, 5000, dtype=np.uint8), 4).tobytes(),
        "runs255": np.repeat(rng.integers(0, 256, 100, dtype=np.uint8), 255).tobytes(),
        "runs256": np.repeat(rng.integers(0, 256, 100, dtype=np.uint8), 256).tobytes(),
        "abc": b"abc" * 5000, "ab": b"ab" * 3000,
        "text": open(os.path.join(ROOT, "SURVEY.md"), "rb").read(),
    }


@pytest.mark.parametrize("i,level", [(1, 1), (2, 2), (3, 3)])
def test_reference_known_answer_vectors(oracle, i, level):
    """src/external/bzip2-1.0.6/Makefile:58-69: `bzip2 -1/-2/-3 < sampleN.ref` must equal sampleN.bz2 byte for byte"""
    z = open(os.path.join(GOLDEN, "sample%d.bz2" % i), "rb").read()
    ref = bz2.decompress(z)
    assert oracle.bz2_compress(ref, level) == z
    rc, back = oracle.bz2_decompress(z, len(ref))
    assert rc == 0 and back == ref


@pytest.mark.parametrize("name", sorted(_cases().keys()))
def test_oracle_matches_libbz2(oracle, name):
    data = _cases()[name]
    for level in (1, 2, 9):
        want = bz2.compress(data, level)
        assert oracle.bz2_compress(data, level) == want, (name, level)
        rc, back = oracle.bz2_decompress(want, len(data))
        assert rc == 0 and back == data


def test_oracle_matches_vendored_bzip2_when_built(oracle):
    """oracle/_ref/libbz2ref.so is the reference's vendored bzip2 1.0.6 compiled as is (oracle/build_ref.py)"""
    so = os.path.join(ROOT, "oracle", "_ref", "libbz2ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    ref = C.CDLL(so)
    ref.BZ2_bzBuffToBuffCompress.argtypes = [C.c_char_p, C.POINTER(C.c_uint), C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_int]
    for name, data in _cases().items():
        for level in (1, 2):
            cap = C.c_uint(len(data) * 2 + 1000); buf = C.create_string_buffer(cap.value)
            assert ref.BZ2_bzBuffToBuffCompress(buf, C.byref(cap), data, len(data), level, 0, 30) == 0
            assert oracle.bz2_compress(data, level) == buf.raw[:cap.value], (name, level)


def test_decoder_rejects_corruption(oracle):
    data = np.random.default_rng(5).poisson(9, 5000).astype(np.uint16).tobytes()
    z = bytearray(bz2.compress(data, 1))
    z[len(z) // 2] ^= 0x10
    rc, _ = oracle.bz2_decompress(bytes(z), len(data))
    assert rc != 0
    assert oracle.bz2_decompress(b"BZx9", 10)[0] == -1


def test_trace_is_consistent(oracle):
    data = np.random.default_rng(2).poisson(5, 9216).astype(np.uint16).tobytes()
    out, blocks = oracle.bz2_compress(data, 1, trace=True)
    assert out == bz2.compress(data, 1) and len(blocks) == 1
    b = blocks[0]
    assert b["mtfv"][-1] == b["n_in_use"] + 1 and len(b["bwt"]) == b["nblock"]
    # the BWT column is a permutation of the block
    assert np.array_equal(np.sort(b["bwt"]), np.sort(b["rle1"]))
