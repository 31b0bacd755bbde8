"""Per-step wall/GPU timing of the device-resident step, to find host-side stalls."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine, batch_from_device
from dropclip_b200.scenes import make_scene
dev = torch.device("cuda", 0)
eng = FusionEngine(dev)
uniq = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda:0", as_torch=True) for i in range(8)]
batch = batch_from_device([uniq[i % 8] for i in range(64)], dev, seg_dtype=torch.int64)
torch.cuda.synchronize()
def step():
    res = eng.fuse_object_level(batch, 0.05, False, True, "max", torch.uint8)
    comp = eng.compact_visibility(batch, res["any_visible"], res["records"], res["rank"], torch.uint8, host_sizes=False)
    return res, comp
for _ in range(3): step()
torch.cuda.synchronize()
gc_log = []
def _gc_cb(phase, info):
    if phase == "start": gc_log.append([time.perf_counter(), info["generation"]])
    else: gc_log[-1][0] = (time.perf_counter() - gc_log[-1][0]) * 1e3
gc.callbacks.append(_gc_cb)
for mode in ("gc on", "gc off"):
    if mode == "gc off":
        gc.collect(); gc.disable()
    for rep in range(3):
        host = []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            n0 = torch.cuda.memory_stats()["num_device_alloc"]; f0 = torch.cuda.memory_stats()["num_device_free"]
            t0 = time.perf_counter(); r = step(); host.append((time.perf_counter() - t0) * 1e3)
            if host[-1] > 5:
                st = torch.cuda.memory_stats()
                print("  slow step: device_alloc +%d device_free +%d, gc events %s" % (st["num_device_alloc"] - n0, st["num_device_free"] - f0, [(round(t, 1), g) for t, g in gc_log[-3:]]))
        b.record(); torch.cuda.synchronize()
        print(mode, f"gpu {a.elapsed_time(b)/10:.2f} ms/step | host launch ms:", " ".join(f"{h:.1f}" for h in host))
