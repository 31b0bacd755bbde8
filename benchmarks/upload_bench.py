"""Host -> device staging strategies for one scene's depth + instance maps (275 MB)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.engine import PinnedStaging

def T():
    torch.cuda.synchronize(); return time.perf_counter()

rng = np.random.default_rng(0)
depths = [rng.random((480, 640), dtype=np.float32) for _ in range(73)]
segs = [rng.integers(0, 21, size=(480, 640)).astype(np.int64) for _ in range(73)]
st = PinnedStaging("cuda")
for gb in (1 << 40, 96 << 20, 48 << 20, 24 << 20, 12 << 20):
    for rep in range(3):
        st.begin(); t0 = T()
        d = st.upload_list(depths, torch.float32, (480, 640), group_bytes=gb); t1 = T()
        s = st.upload_list(segs, torch.int64, (480, 640), group_bytes=gb); t2 = T()
        st.end()
    print(f"group {gb>>20:8d} MB: depth {1e3*(t1-t0):6.2f} ms  seg {1e3*(t2-t1):6.2f} ms  total {1e3*(t2-t0):6.2f}")
# pageable direct
for rep in range(3):
    t0 = T(); d = torch.from_numpy(np.stack(depths)).cuda(); s = torch.from_numpy(np.stack(segs)).cuda(); t1 = T()
print(f"np.stack + pageable .cuda(): {1e3*(t1-t0):.2f} ms")
for rep in range(3):
    t0 = T(); ds = [torch.from_numpy(x).cuda(non_blocking=True) for x in depths]; ss = [torch.from_numpy(x).cuda(non_blocking=True) for x in segs]; t1 = T()
print(f"per-array pageable .cuda(): {1e3*(t1-t0):.2f} ms")
# cudaHostRegister in place
import ctypes
cudart = ctypes.CDLL("libcudart.so.12")
for rep in range(2):
    t0 = T()
    for x in segs:
        cudart.cudaHostRegister(ctypes.c_void_p(x.ctypes.data), ctypes.c_size_t(x.nbytes), 0)
    t1 = T()
    outs = [torch.from_numpy(x).cuda(non_blocking=True) for x in segs]
    t2 = T()
    for x in segs:
        cudart.cudaHostUnregister(ctypes.c_void_p(x.ctypes.data))
    t3 = T()
    print(f"hostRegister 73 seg maps: register {1e3*(t1-t0):.2f} ms, H2D {1e3*(t2-t1):.2f} ms, unregister {1e3*(t3-t2):.2f} ms")
