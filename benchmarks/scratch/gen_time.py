import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from dropclip_b200.scenes import make_scene
torch.cuda.synchronize()
for i in range(6):
    t0 = time.perf_counter()
    sc = make_scene(5000 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda", as_torch=True)
    torch.cuda.synchronize()
    print("scene", i, time.perf_counter() - t0, "s")
print(torch.cuda.max_memory_allocated() / 1e9, "GB peak")
