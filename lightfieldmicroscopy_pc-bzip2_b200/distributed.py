"""Multi-process (one rank per GPU) writer / reader for .lfm stacks.

KLB blocks are independent bzip2 streams and block ids are x fastest, z/c/t slowest
(reference: src/klb_imageIO.cpp:133-140), so a contiguous range of z-slabs is a contiguous block-id range and a
contiguous byte range of the payload. Each rank therefore compresses its own slab range on its own GPU with NO
collective on the data path; the only exchange is tiny and host-driven:

  * rank 0 runs the predictor mode selection on frame 0 and broadcasts the 3-bit result
    (src/klb_imageIO.cpp:2316-2377 selects on frame 0 only),
  * the per-block compressed sizes are all-reduced (disjoint supports) and prefix-summed on the host into
    header.blockOffset[] (the job of blockWriter, src/klb_imageIO.cpp:1145-1225),
  * rank 0 writes the 320-byte header + the table and sizes the file; every rank streams its payload from its GPU to its
    byte offset (lfmShardWritePayload: pinned ring + pwrite, no payload-sized host copy).

Reading is the mirror image: every rank reads the header, takes its slab range and decodes it with readKLBroiInPlace, which
fetches only the byte range of those slabs.

`torch.distributed` is only plumbing (gloo or nccl; the tensors are 8 bytes per block and live on the GPU under nccl).
"""
import math
import os
import struct

import numpy as np

HEADER_FIXED = 320


def slab_partition(xyzct, block_size, world):
    """z-slab ranges per rank: [(first_slab, n_slabs, z0, z1)], slabs aligned to lcm(blockSize[2], 2) frames so that
    video stacks (odd frames predicted from the even frame before them) never straddle ranks."""
    x, y, z, c, t = xyzct
    bz = min(block_size[2], z)
    assert c == 1 and t == 1, "sharded writer handles xyz stacks (present video stacks as z = frames)"
    nbz = int(math.ceil(np.float32(z) / np.float32(bz)))
    group = 1 if bz % 2 == 0 else 2          # slabs per indivisible group
    ngroups = (nbz + group - 1) // group
    out = []
    for r in range(world):
        g0 = ngroups * r // world; g1 = ngroups * (r + 1) // world
        s0 = min(nbz, g0 * group); s1 = min(nbz, g1 * group)
        out.append((s0, s1 - s0, min(z, s0 * bz), min(z, s1 * bz)))
    return out


def pack_header(header_version, nnum, xyzct, block_size, block_offset, pixel_size=(1.0,) * 5, metadata=b""):
    """the reference's on-disk header (src/klb_imageHeader.cpp:164-176)"""
    h = struct.pack("<BB5I5fBB", header_version, nnum, *xyzct, *pixel_size, 1, 1)
    h += metadata.ljust(256, b"\0")[:256]
    h += struct.pack("<5I", *block_size)
    assert len(h) == HEADER_FIXED
    return h + np.asarray(block_offset, dtype="<u8").tobytes()


def parse_header(filename):
    """(headerVersion, Nnum, xyzct, blockSize) of a .lfm file"""
    with open(filename, "rb") as f:
        b = f.read(HEADER_FIXED)
    hv, nnum = b[0], b[1]
    xyzct = struct.unpack_from("<5I", b, 2)
    bs = struct.unpack_from("<5I", b, 300)
    return hv, nnum, tuple(xyzct), tuple(bs)


class GpuShardBackend:
    """the product path: this rank's GPU engine through the C ABI (include/lfm_b200.h)"""

    def __init__(self, nnum, block_size, way):
        import importlib
        self.L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
        self.nnum, self.bs, self.way = nnum, tuple(block_size), way

    def select_mode(self, frame0):
        bs = (self.bs[0], self.bs[1], 1, 1, 1)
        shv, _, _ = self.L.shard_compress(np.ascontiguousarray(frame0[None]), 0, nnum=self.nnum, block_size=bs, way=self.way)
        return shv & 0x7F

    def compress(self, frames, header_version):
        """-> (uint32 sizes of the local blocks, payload bytes); the payload stays on the GPU"""
        _, sizes, pb = self.L.shard_compress(np.ascontiguousarray(frames), header_version, nnum=self.nnum, block_size=self.bs, way=self.way)
        return sizes, pb

    def write_header(self, filename, xyzct, bs, stored_hv, nnum, block_offset):
        self.L.write_header(filename, xyzct, bs, stored_hv, nnum, block_offset)       # lfmWriteHeader: header + table, file sized

    def write_payload(self, filename, file_offset):
        self.L.shard_write_payload(filename, file_offset)

    def read_frames(self, filename, xyzct, z0, z1):
        return self.L.read_roi(filename, (0, 0, z0, 0, 0), (xyzct[0] - 1, xyzct[1] - 1, z1 - 1, 0, 0), way=self.way)[0, 0]


def _exchange_device(dist):
    """tensors of the size exchange live where the backend can reduce them"""
    if dist.get_backend() == "nccl":
        import torch
        return torch.device("cuda", torch.cuda.current_device())
    return "cpu"


def write_stack_sharded(local_frames, xyzct, filename, header_version=0, nnum=13, block_size=(96, 96, 8, 1, 1), way=0,
                        backend=None, dist=None):
    """Collective call. `local_frames`: this rank's frames [z0:z1] (uint16 [n, y, x]) as given by slab_partition().
    backend: object with select_mode / compress / write_payload (default: the GPU engine). Returns the stored headerVersion."""
    if dist is None:
        import torch.distributed as dist
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    x, y, z, c, t = xyzct
    bs = tuple(min(b, d) for b, d in zip(block_size, xyzct))
    parts = slab_partition(xyzct, bs, world)
    s0, ns, z0, z1 = parts[rank]
    if backend is None:
        backend = GpuShardBackend(nnum, bs, way)
    dev = _exchange_device(dist)
    # ---- predictor: forced, or selected on frame 0 by the rank that owns it and broadcast
    video = header_version & 0x80
    k = torch.zeros(1, dtype=torch.int32, device=dev)
    if (header_version & 0x7F) < 8:
        if rank == 0:
            k[0] = int(backend.select_mode(np.ascontiguousarray(local_frames[0])))
        dist.broadcast(k, src=0)
    else:
        k[0] = header_version & 0x77 & 0x7F
    k = int(k[0])
    stored_hv = video | k
    # ---- my blocks (no collective)
    nb_xy = int(math.ceil(np.float32(x) / np.float32(bs[0]))) * int(math.ceil(np.float32(y) / np.float32(bs[1])))
    if ns > 0:
        sizes, _ = backend.compress(local_frames, video | (8 + k))
        assert sizes.size == ns * nb_xy
    else:
        sizes = np.zeros(0, np.uint32)
    # ---- exchange of the block sizes, host prefix sum -> blockOffset
    nbz = int(math.ceil(np.float32(z) / np.float32(bs[2])))
    nb_total = nbz * nb_xy
    mine = torch.zeros(nb_total, dtype=torch.int64)
    mine[s0 * nb_xy:(s0 + ns) * nb_xy] = torch.from_numpy(sizes.astype(np.int64))
    mine = mine.to(dev)
    dist.all_reduce(mine)                      # disjoint supports: the sum is the concatenation
    block_offset = np.cumsum(mine.cpu().numpy()).astype(np.uint64)
    my_off = int(block_offset[s0 * nb_xy - 1]) if s0 * nb_xy > 0 else 0
    if rank == 0:
        if hasattr(backend, "write_header"):
            backend.write_header(filename, xyzct, bs, stored_hv, nnum, block_offset)
        else:
            with open(filename, "wb") as f:
                f.write(pack_header(stored_hv, nnum, xyzct, bs, block_offset))
                f.truncate(HEADER_FIXED + 8 * nb_total + int(block_offset[-1]))
    dist.barrier()
    if ns > 0:
        backend.write_payload(filename, HEADER_FIXED + 8 * nb_total + my_off)
    dist.barrier()
    return stored_hv


def read_stack_sharded(filename, way=0, backend=None, dist=None):
    """Collective in name only (no exchange): every rank decodes the z range slab_partition() gives it.
    Returns (z0, z1, frames[z1 - z0, y, x])."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    hv, nnum, xyzct, bs = parse_header(filename)
    s0, ns, z0, z1 = slab_partition(xyzct, bs, world)[rank]
    if backend is None:
        backend = GpuShardBackend(nnum, bs, way)
    if ns == 0:
        return z0, z1, np.zeros((0, xyzct[1], xyzct[0]), np.uint16)
    return z0, z1, backend.read_frames(filename, xyzct, z0, z1)
