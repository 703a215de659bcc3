"""CPU-side tests: the C-ABI library loads and exports what include/*.h declares, host logic, error behaviour
without a GPU, and the N>1 (one rank per GPU) host protocol on gloo with world_size 2."""
import ctypes as C
import os
import re
import struct
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_img_tif, has_cuda, lf_synth


@pytest.fixture(scope="module")
def L():
    import lfm_b200
    return lfm_b200


def _declared_functions():
    names = []
    for h in ("klb_Cwrapper.h", "lfm_b200.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(?:int|void\*|uint64_t|const char\*)\s+\**(\w+)\s*\(", txt)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(L):
    names = _declared_functions()
    assert len(names) >= 19 and "writeKLBstack" in names and "lfmCompressDevice" in names
    for n in names:
        assert hasattr(L.lib, n), "include/*.h declares %s but liblfm_b200.so does not export it" % n
    assert set(L.EXPORTS) <= set(names)


def test_threaded_host_copy_copies_every_byte(L):
    """the staging copies split a buffer over the copy pool in 64-byte aligned parts: every length must arrive whole (a length whose
    per-thread share is a multiple of 64 once lost its last n mod threads bytes: the tail of a 241 MB payload, found at configs[4])"""
    import ctypes as C
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (3 << 20) + 4096, dtype=np.uint8)
    dst = np.zeros_like(src)
    threads = [2, 3, 4, 5, 6, 7, 8]
    sizes = {(1 << 20) + k for k in range(0, 9)} | {512 * 1024, 512 * 1024 + 1, (3 << 20) + 77}
    sizes |= {64 * t * m + r for t in threads for m in (4096, 4097) for r in range(1, t)}     # floor(n / t) a multiple of 64, n mod t != 0
    for n in sorted(sizes):
        dst[:] = 0
        assert L.lib.lfmDebugParMemcpy(C.c_void_p(dst.ctypes.data), C.c_void_p(src.ctypes.data), C.c_uint64(n)) == 0
        assert np.array_equal(dst[:n], src[:n]), "par_memcpy dropped bytes at n=%d" % n
        assert not dst[n:n + 64].any()


def test_way_setter(L):
    prev = L.lib.lfmGetPredictorWay()
    assert L.lib.lfmSetPredictorWay(2) == prev and L.lib.lfmGetPredictorWay() == 2
    assert L.lib.lfmSetPredictorWay(7) == -1 and L.lib.lfmGetPredictorWay() == 2
    L.lib.lfmSetPredictorWay(prev)


def test_num_blocks_matches_reference_rule(L):
    for shape, bs, want in [((29, 151, 101), None, 16), ((1, 2048, 2048), None, 484), ((101, 2048, 2048), None, 6292),
                            ((200, 4096, 4096), None, 46225), ((5, 10, 7), (4, 4, 2, 1, 1), 2 * 3 * 3)]:
        b = L._u32x5(*bs) if bs else None
        assert L.lib.lfmNumBlocks(L._xyzct(shape), C.cast(b, C.c_void_p) if b is not None else None) == want


def test_header_reader_on_oracle_file(L, oracle, tmp_path):
    a = golden_img_tif()
    fn = str(tmp_path / "o.lfm")
    rc, shv = oracle.write(a, fn, 8 + 5, 13, 0)
    assert rc == 0
    h = L.read_header(fn)
    assert h["xyzct"] == [101, 151, 29, 1, 1] and h["blockSize"] == [96, 96, 8, 1, 1] and h["headerVersion"] == 5 and h["Nnum"] == 13
    assert h["dataType"] == 1 and h["compressionType"] == 1 and h["pixelSize"] == [1.0] * 5
    with pytest.raises(L.LfmError):
        L.read_header(str(tmp_path / "missing.lfm"))


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(L, tmp_path):
    """the product path must fail loudly without a CUDA device (code 6), never fall back to a CPU codec"""
    a = lf_synth((2, 40, 40), 13)
    with pytest.raises(L.LfmError) as ei:
        L.compress_to_bytes(a)
    assert ei.value.code == 6
    xyzct = L._u32x5(40, 40, 2, 1, 1)
    assert L.lib.writeKLBstack(a.ctypes.data, os.fsencode(str(tmp_path / "x.lfm")), xyzct, 1, -1, None, None, 1, None) == 6
    assert L.lib.writeKLBstack(a.ctypes.data, b"/nonexistent_dir/x.lfm", xyzct, 1, -1, None, None, 1, None) == 5


def test_slab_partition_is_a_partition():
    sys.path.insert(0, os.path.join(ROOT, "lightfieldmicroscopy_pc-bzip2_b200"))
    import distributed as D
    for z, bz, world in [(101, 8, 8), (1000, 1, 8), (29, 8, 2), (16, 3, 4), (5, 8, 4)]:
        parts = D.slab_partition((64, 64, z, 1, 1), (32, 32, bz, 1, 1), world)
        assert parts[0][2] == 0 and parts[-1][3] == z
        for (a, b) in zip(parts, parts[1:]):
            assert a[3] == b[2] and a[0] + a[1] == b[0]
        for p in parts:
            assert p[2] % 2 == 0 or p[2] == z          # video stacks: every shard starts on an even frame


class _OracleBackend:
    """CPU stand-in for the GPU engine (tests only): the oracle writes the slab as a stack of its own, the payload is cut out"""

    def __init__(self, ora, tmpdir, rank, bs):
        self.ora, self.fn, self.bs, self.payload = ora, os.path.join(tmpdir, "slab_%d.lfm" % rank), bs, b""

    def select_mode(self, f0):
        return self.ora.select(f0, 13, 0)[0]

    def compress(self, frames, hv):
        rc, _ = self.ora.write(frames, self.fn, hv, 13, 0, block_size=self.bs)
        assert rc == 0
        blob = open(self.fn, "rb").read()
        nb = int(np.prod([-(-d // b) for d, b in zip((frames.shape[2], frames.shape[1], frames.shape[0]), self.bs[:3])]))
        ends = np.frombuffer(blob[320:320 + 8 * nb], "<u8")
        self.payload = blob[320 + 8 * nb:]
        return np.diff(np.concatenate([[0], ends])).astype(np.uint32), len(self.payload)

    def write_payload(self, filename, off):
        fd = os.open(filename, os.O_WRONLY)
        try:
            os.pwrite(fd, self.payload, off)
        finally:
            os.close(fd)

    def read_frames(self, filename, xyzct, z0, z1):
        rc, out = self.ora.read(filename, (xyzct[2], xyzct[1], xyzct[0]), 0)
        assert rc == 0
        return out[z0:z1]


def _worker(rank, world, port, tmpdir, gpu):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "lightfieldmicroscopy_pc-bzip2_b200"))
    import distributed as D
    from conftest import Oracle, lf_synth
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ora = Oracle()
    a = lf_synth((21, 70, 90), 13, seed=5)
    xyzct = (90, 70, 21, 1, 1); bs = (32, 32, 4, 1, 1)
    backend = None
    if gpu:                                   # the product path: every rank drives the GPU engine through the C ABI
        import torch
        import lfm_b200
        lfm_b200.set_devices(rank % torch.cuda.device_count(), 1)
    else:
        backend = _OracleBackend(ora, tmpdir, rank, bs)
    for hv, name in ((0, "auto"), (0x80 | 12, "video4")):
        z0, z1 = D.slab_partition(xyzct, bs, world)[rank][2:]
        out = os.path.join(tmpdir, "sharded_%s.lfm" % name)
        D.write_stack_sharded(a[z0:z1], xyzct, out, header_version=hv, nnum=13, block_size=bs, way=0, backend=backend, dist=dist)
        if rank == 0:
            whole = os.path.join(tmpdir, "whole_%s.lfm" % name)
            rc, _ = ora.write(a, whole, hv, 13, 0, block_size=bs)
            assert rc == 0
            assert open(out, "rb").read() == open(whole, "rb").read(), "sharded file differs from the single-process file (%s)" % name
        r0, r1, fr = D.read_stack_sharded(out, way=0, backend=backend, dist=dist)
        assert (r0, r1) == (z0, z1) and np.array_equal(fr, a[z0:z1]), "sharded read back (%s)" % name
    dist.destroy_process_group()


def test_two_rank_sharded_writer_on_gloo(tmp_path):
    """world_size 2 on CPU: per-rank slabs, size exchange + host prefix sum, pwrite at offsets == single-writer file; sharded read"""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path), False), nprocs=2, join=True)


@pytest.mark.gpu
def test_two_rank_sharded_writer_gpu_engine(tmp_path):
    """the same protocol with the GPU engine behind every rank (ranks share the GPUs present: world_size 2 also on a 1-GPU box):
    lfmShardCompress -> size exchange -> lfmWriteHeader / lfmShardWritePayload at offsets -> file == the oracle's single-writer
    file; read_stack_sharded (readKLBroiInPlace per rank) returns every rank's frames"""
    import torch.multiprocessing as mp
    port = 29500 + ((os.getpid() + 17) % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path), True), nprocs=2, join=True)
