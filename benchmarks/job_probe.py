"""Per-batch device and host time of the distinct-scene job loop of bench.py (fresh batch per pass)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine, batch_from_device
from dropclip_b200.scenes import make_scene
dev = torch.device("cuda", 0)
eng = FusionEngine(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
def step(b):
    res = eng.fuse_object_level(b, 0.05, False, True, "max", torch.uint8, join=False)
    comp = eng.compact_visibility(b, res["any_visible"], res["records"], res["rank"], torch.uint8, host_sizes=False)
    res["join"]()
    return res, comp
for overlap in (True, False, True):
    eng.overlap = overlap
    torch.cuda.empty_cache()
    for bi in range(4):
        scs = [make_scene(100_000 + bi * n + i, n_views=73, n_points=100_000, n_objects=21, device="cuda:0", as_torch=True) for i in range(n)]
        jb = batch_from_device(scs, dev, seg_dtype=torch.int64)
        del scs
        torch.cuda.synchronize()
        st0 = torch.cuda.memory_stats()["num_device_alloc"]
        j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.profile = {}
        t0 = time.perf_counter()
        j0.record()
        jr = step(jb)
        j1.record()
        host = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        prof = eng.profile_ms(); eng.profile = None
        print(f"overlap={int(overlap)} batch {bi}: device {j0.elapsed_time(j1):.2f} ms, host enqueue {host:.2f} ms, cudaMallocs {torch.cuda.memory_stats()['num_device_alloc'] - st0}  "
              + " ".join(f"{k}={v:.2f}" for k, v in prof.items()), flush=True)
        del jb, jr
