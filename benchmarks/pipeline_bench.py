"""FusionPipeline throughput at the configs[1] sizes: pre-filled pinned slots cycled through H2D -> fuse -> D2H."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200.pipeline import FusionPipeline

B = int(os.environ.get("B", "4")); NS = int(os.environ.get("SLOTS", "12")); N = int(os.environ.get("SCENES", "96"))
dev = torch.device("cuda", 0)
scs = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda") for i in range(NS)]
pipe = FusionPipeline(scs[0].intrinsic, device=dev, batch_scenes=B, n_slots=NS)
slots = []
for sc in scs:
    s = pipe.acquire()
    s.fill(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings)
    slots.append(s)
for s in slots:
    pipe._free.put(s)
# raw PCIe rate: one pinned->device copy of a slot's depth block
d = torch.empty_like(slots[0].t_depths, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(slots[0].t_depths, non_blocking=True)
torch.cuda.synchronize()
pcie = 10 * slots[0].t_depths.numel() * 4 / (time.perf_counter() - t0) / 1e9
print("pinned H2D GB/s", pcie)
n_done = [0]
def consume():
    for r in pipe.results():
        n_done[0] += 1
t = threading.Thread(target=consume); t.start()
for rep in range(2):
    torch.cuda.synchronize()
    n0, b0 = n_done[0], pipe.h2d_bytes
    t0 = time.perf_counter()
    for i in range(N):
        s = pipe.acquire()
        pipe.submit(s, tag=i)
    while n_done[0] < n0 + N:
        time.sleep(0.0005)
        if pipe._error is not None or time.perf_counter() - t0 > 60:
            raise SystemExit(f"pipeline stopped: {pipe._error!r}")
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {N / dt:.1f} scenes/s, H2D {(pipe.h2d_bytes - b0) / dt / 1e9:.1f} GB/s ({(pipe.h2d_bytes - b0) / N / 1e6:.1f} MB/scene), ms/scene {dt / N * 1e3:.2f}")
pipe.finish(); t.join(); pipe.close()
