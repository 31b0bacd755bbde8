"""Where does upload_list spend its host time? helper calls vs H2D enqueue vs waits, for a fresh (cold) scene each time."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200 import _lib
from dropclip_b200.engine import PinnedStaging

lib = _lib.load()
V, H, W = 73, 480, 640
rng = np.random.default_rng(0)
scenes = [([rng.random((H, W), dtype=np.float32) for _ in range(V)], [rng.integers(0, 21, size=(H, W), dtype=np.int64) for _ in range(V)])
          for _ in range(4)]
st = PinnedStaging("cuda")
for gmb in (8, 32, 128, 1024):
    for rep in range(2):
        tot = []
        for depths, segs in scenes:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st.begin()
            d = st.upload_list(depths, torch.float32, (H, W), group_bytes=gmb << 20)
            t1 = time.perf_counter()
            s, ok = st.upload_list(segs, torch.int64, (H, W), group_bytes=gmb << 20, narrow_to_u8=True)
            t2 = time.perf_counter()
            st.end()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            tot.append((t1 - t0, t2 - t1, t3 - t2))
    a = np.array(tot) * 1e3
    print(f"group {gmb:5d} MB: depth host {a[:,0].mean():5.2f} ms | seg host {a[:,1].mean():5.2f} ms | drain {a[:,2].mean():5.2f} ms | total {a.sum(1).mean():5.2f} ms")
