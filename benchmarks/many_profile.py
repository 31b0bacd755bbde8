"""Where does fuse_many spend its time? Times the staging thread and the finish step separately."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200 import feature_fusion as ff
host = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda") for i in range(4)]
M = ff.MultiviewFeatureFusion(host[0].intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
args = [(s.points, s.colors, s.labels, s.depths, s.seg_masks, s.camera_poses, s.mv_features, s.query_embeddings) for s in host]
log = []
orig_stage, orig_finish = M._stage_obj, M._finish_obj
def stage(*a, **k):
    t0 = time.perf_counter(); r = orig_stage(*a, **k); log.append(("stage", threading.current_thread().name, (time.perf_counter() - t0) * 1e3)); return r
def finish(*a, **k):
    t0 = time.perf_counter(); r = orig_finish(*a, **k); log.append(("finish", threading.current_thread().name, (time.perf_counter() - t0) * 1e3)); return r
M._stage_obj, M._finish_obj = stage, finish
def stats():
    d = torch.cuda.memory_stats()
    h = torch.cuda.host_memory_stats() if hasattr(torch.cuda, "host_memory_stats") else {}
    return d["num_device_alloc"], h.get("num_host_alloc", -1)
for rep in range(8):
    log.clear()
    s0 = stats()
    t0 = time.perf_counter()
    outs = [(f.cpu(), w.cpu(), vis) for (f, w, vis), _ in M.fuse_many(args, return_obj=True, device="cuda")]
    dt = (time.perf_counter() - t0) * 1e3
    s1 = stats()
    print(f"rep {rep}: {dt / 4:.1f} ms/scene | device allocs +{s1[0]-s0[0]} host allocs +{s1[1]-s0[1]} |", " ".join(f"{k}:{t:.1f}" for k, _, t in log))
