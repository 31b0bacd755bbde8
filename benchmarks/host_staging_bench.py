"""Host staging helpers: GB/s of the multi-threaded gather-copy and int64->uint8 narrowing vs thread count."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200 import _lib
lib = _lib.load()
V, H, W = 73, 480, 640
rng = np.random.default_rng(0)
depths = [rng.random((H, W), dtype=np.float32) for _ in range(V)]
segs = [rng.integers(0, 21, size=(H, W), dtype=np.int64) for _ in range(V)]
pd = torch.empty((V, H, W), dtype=torch.float32, pin_memory=True)
ps = torch.empty((V, H, W), dtype=torch.uint8, pin_memory=True)
sd = (ctypes.c_void_p * V)(*[a.ctypes.data for a in depths])
ss = (ctypes.c_void_p * V)(*[a.ctypes.data for a in segs])
bad = ctypes.c_int(0)
for nt in (1, 2, 4, 6, 8, 12, 16):
    for _ in range(2):
        lib.dc_host_gather_copy(sd, V, H * W * 4, ctypes.c_void_p(pd.data_ptr()), nt)
        lib.dc_host_gather_narrow_i64_u8(ss, V, H * W, ctypes.c_void_p(ps.data_ptr()), nt, ctypes.byref(bad))
    t0 = time.perf_counter()
    for _ in range(5):
        lib.dc_host_gather_copy(sd, V, H * W * 4, ctypes.c_void_p(pd.data_ptr()), nt)
    t1 = time.perf_counter()
    for _ in range(5):
        lib.dc_host_gather_narrow_i64_u8(ss, V, H * W, ctypes.c_void_p(ps.data_ptr()), nt, ctypes.byref(bad))
    t2 = time.perf_counter()
    print(f"threads {nt:2d}: copy 90MB {(t1-t0)/5*1e3:6.2f} ms ({0.0897/((t1-t0)/5):5.1f} GB/s)   narrow 179MB {(t2-t1)/5*1e3:6.2f} ms ({0.1794/((t2-t1)/5):5.1f} GB/s read)")
assert bad.value == 0 and np.array_equal(ps[5].numpy(), segs[5].astype(np.uint8)) and np.array_equal(pd[7].numpy(), depths[7])

# comparison points on the same box: torch's own (multi-threaded) copy of the stacked array into pinned memory, and the
# same gather when the sources were evicted from the caches in between (what fuse() sees for a fresh scene)
stacked = torch.from_numpy(np.stack(depths))
for _ in range(2):
    pd.copy_(stacked)
t0 = time.perf_counter()
for _ in range(5):
    pd.copy_(stacked)
print(f"torch copy_ 90MB pageable->pinned {(time.perf_counter()-t0)/5*1e3:6.2f} ms")
evict = np.zeros(512 << 20, dtype=np.uint8)
for nt in (8, 16):
    ts, tn = [], []
    for _ in range(4):
        evict += 1  # 512 MB read+write: the sources leave the caches
        t0 = time.perf_counter()
        lib.dc_host_gather_copy(sd, V, H * W * 4, ctypes.c_void_p(pd.data_ptr()), nt)
        t1 = time.perf_counter()
        lib.dc_host_gather_narrow_i64_u8(ss, V, H * W, ctypes.c_void_p(ps.data_ptr()), nt, ctypes.byref(bad))
        t2 = time.perf_counter()
        ts.append(t1 - t0); tn.append(t2 - t1)
    print(f"cold sources, threads {nt:2d}: copy 90MB {min(ts)*1e3:6.2f} ms   narrow 179MB {min(tn)*1e3:6.2f} ms")
evict += 1
t0 = time.perf_counter()
pd.copy_(stacked)
print(f"cold torch copy_ 90MB {(time.perf_counter()-t0)*1e3:6.2f} ms")
