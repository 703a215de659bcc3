"""Multi-process (one rank per GPU) writer / reader for .lfm stacks.

KLB blocks are independent bzip2 streams and block ids are x fastest, z/c/t slowest
(reference: src/klb_imageIO.cpp:133-140), so a contiguous range of z-slabs is a contiguous block-id range and a
contiguous byte range of the payload. Each rank therefore compresses its own slab range on its own GPU with NO
collective on the data path; the only exchange is host-side and tiny:

  * rank 0 runs the predictor mode selection on frame 0 and broadcasts the 3-bit result
    (src/klb_imageIO.cpp:2316-2377 selects on frame 0 only),
  * the per-block compressed sizes are all-gathered and prefix-summed on the host into header.blockOffset[]
    (the job of blockWriter, src/klb_imageIO.cpp:1145-1225),
  * every rank pwrite()s its payload at its byte offset; rank 0 writes the 320-byte header + the table.

`torch.distributed` is only plumbing here (gloo or nccl; the tensors are a few bytes per block).
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


def write_stack_sharded(local_frames, xyzct, filename, header_version=0, nnum=13, block_size=(96, 96, 8, 1, 1), way=0,
                        compress_slab=None, select_mode=None, dist=None):
    """Collective call. `local_frames`: this rank's frames [z0:z1] (uint16 [n, y, x]) as given by slab_partition().
    compress_slab(frames, forced_header_version) -> (.lfm bytes of the slab written as its own stack);
    select_mode(frame0) -> predictor 0..7. Both default to the GPU engine through the C ABI."""
    if dist is None:
        import torch.distributed as dist
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    x, y, z, c, t = xyzct
    bs = tuple(min(b, d) for b, d in zip(block_size, xyzct))
    parts = slab_partition(xyzct, bs, world)
    s0, ns, z0, z1 = parts[rank]
    if compress_slab is None or select_mode is None:
        import importlib
        L = importlib.import_module("lightfieldmicroscopy_pc-bzip2_b200")
        if compress_slab is None:
            compress_slab = lambda fr, hv: L.compress_to_bytes(fr, header_version=hv, nnum=nnum, block_size=bs, way=way)
        if select_mode is None:
            def select_mode(f0):
                blob = L.compress_to_bytes(f0[None], header_version=0, nnum=nnum, block_size=(bs[0], bs[1], 1, 1, 1), way=way)
                return blob[0] & 0x7F
    # ---- predictor: forced, or selected on frame 0 by the rank that owns it and broadcast
    video = header_version & 0x80
    k = torch.zeros(1, dtype=torch.int32)
    if (header_version & 0x7F) < 8:
        if rank == 0:
            k[0] = int(select_mode(np.ascontiguousarray(local_frames[0])))
        dist.broadcast(k, src=0)
    else:
        k[0] = header_version & 0x77 & 0x7F
    k = int(k[0])
    stored_hv = video | k
    # ---- my blocks (no collective)
    nb_xy = int(math.ceil(np.float32(x) / np.float32(bs[0]))) * int(math.ceil(np.float32(y) / np.float32(bs[1])))
    if ns > 0:
        blob = compress_slab(np.ascontiguousarray(local_frames), video | (8 + k))
        nb_local = ns * nb_xy
        ends = np.frombuffer(blob[HEADER_FIXED:HEADER_FIXED + 8 * nb_local], dtype="<u8")
        payload = blob[HEADER_FIXED + 8 * nb_local:]
        sizes = np.diff(np.concatenate([[0], ends])).astype(np.int64)
    else:
        payload = b""; sizes = np.zeros(0, np.int64)
    # ---- host-side exchange of the block sizes, prefix sum -> blockOffset
    nbz = int(math.ceil(np.float32(z) / np.float32(bs[2])))
    nb_total = nbz * nb_xy
    mine = torch.zeros(nb_total, dtype=torch.int64)
    mine[s0 * nb_xy:(s0 + ns) * nb_xy] = torch.from_numpy(sizes)
    dist.all_reduce(mine)                      # disjoint supports: the sum is the concatenation
    block_offset = np.cumsum(mine.numpy()).astype(np.uint64)
    my_off = int(block_offset[s0 * nb_xy - 1]) if s0 > 0 and nb_xy * s0 > 0 else 0
    if rank == 0:
        with open(filename, "wb") as f:
            f.write(pack_header(stored_hv, nnum, xyzct, bs, block_offset))
            f.truncate(HEADER_FIXED + 8 * nb_total + int(block_offset[-1]))
    dist.barrier()
    fd = os.open(filename, os.O_WRONLY)
    try:
        os.pwrite(fd, payload, HEADER_FIXED + 8 * nb_total + my_off)
    finally:
        os.close(fd)
    dist.barrier()
    return stored_hv
