"""Host -> device upload of a large pageable array: staging-copy and end-to-end rates (development tool)."""
import os, sys, time
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffxpy_b200 import device as dxd
n = 276_000_000
a = np.random.default_rng(0).integers(0, 100000, n, dtype=np.int32)
print("cpu count", os.cpu_count(), "torch threads", torch.get_num_threads())
chunk = 8 << 20
pin = torch.empty(chunk, dtype=torch.int32).pin_memory()
t = torch.from_numpy(a)
for nt in (1, 4, 8, 16):
    torch.set_num_threads(nt)
    t0 = time.perf_counter()
    for i in range(16):
        pin.copy_(t[i * chunk:(i + 1) * chunk])
    dt = time.perf_counter() - t0
    print("torch copy_ threads %2d: %.1f GB/s" % (nt, 16 * chunk * 4 / dt / 1e9))
pn = pin.numpy()
for nt in (2, 4, 8, 16):
    pool = ThreadPoolExecutor(nt)
    def cp(i, k):
        s = chunk // nt
        np.copyto(pn[k * s:(k + 1) * s], a[i * chunk + k * s:i * chunk + (k + 1) * s])
    t0 = time.perf_counter()
    for i in range(16):
        list(pool.map(lambda k: cp(i, k), range(nt)))
    dt = time.perf_counter() - t0
    print("numpy copyto pool %2d: %.1f GB/s" % (nt, 16 * chunk * 4 / dt / 1e9))
dev = torch.device("cuda", 0)
torch.set_num_threads(os.cpu_count() or 8)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = dxd.upload(a, dev, torch.int32); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("upload() int32: %.1f ms, %.1f GB/s, equal %s" % (dt * 1e3, n * 4 / dt / 1e9, bool(torch.equal(d.cpu(), torch.from_numpy(a))) if rep == 3 else "-"))
a64 = a[:100_000_000].astype(np.float64) * 0.5
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = dxd.upload(a64, dev, torch.float32); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("upload() float64 -> float32: %.1f ms, %.1f GB/s of output, equal %s" % (dt * 1e3, a64.size * 4 / dt / 1e9, bool(torch.equal(d.cpu(), torch.from_numpy(a64.astype(np.float32))))))
