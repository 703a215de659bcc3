"""Host <-> device copy floor on this box: pinned DMA both ways, single-thread memcpy pageable <-> pinned, /dev/shm write / read of a payload.
   python tools/pcie_probe.py"""
import time, os, numpy as np, torch
for mb in (4.2, 8.4, 64):
    n = int(mb * 1e6)
    pin = torch.empty(n, dtype=torch.uint8).pin_memory()
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    page = np.random.randint(0, 255, n, dtype=np.uint8)
    for _ in range(3): dev.copy_(pin, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): dev.copy_(pin, non_blocking=True)
    torch.cuda.synchronize(); h2d = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(20): pin.copy_(dev, non_blocking=True)
    torch.cuda.synchronize(); d2h = (time.perf_counter() - t0) / 20
    pn = pin.numpy()
    t0 = time.perf_counter()
    for _ in range(20): np.copyto(pn, page)
    mc = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(10): dev.copy_(torch.from_numpy(page)); torch.cuda.synchronize()
    pg = (time.perf_counter() - t0) / 10
    fn = "/dev/shm/pcie_probe.bin"
    fd = os.open(fn, os.O_RDWR | os.O_CREAT, 0o644)
    os.pwrite(fd, pn.tobytes(), 0)
    buf = pn.tobytes()
    t0 = time.perf_counter()
    for _ in range(10): os.pwrite(fd, buf, 0)
    fw = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(10): os.preadv(fd, [pn], 0)
    fr = (time.perf_counter() - t0) / 10
    os.close(fd); os.remove(fn)
    print("%.1f MB: pinned H2D %.3f ms (%.1f GB/s) | pinned D2H %.3f ms (%.1f GB/s) | memcpy pageable->pinned 1 thread %.3f ms (%.1f GB/s) | torch pageable H2D %.3f ms | /dev/shm rewrite %.3f ms, read %.3f ms"
          % (mb, h2d * 1e3, n / h2d / 1e9, d2h * 1e3, n / d2h / 1e9, mc * 1e3, n / mc / 1e9, pg * 1e3, fw * 1e3, fr * 1e3))
