"""bench.py's e2e loop in isolation: per-call times of fuse() over 4 distinct scenes, results kept like bench does."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200.feature_fusion import MultiviewFeatureFusion
host = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda") for i in range(4)]
M = MultiviewFeatureFusion(host[0].intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
outs = None
for step in range(6):
    ts = []
    new = []
    for s in host:
        t0 = time.perf_counter()
        (f, w, vis), (p, c, l) = M.fuse(s.points, s.colors, s.labels, s.depths, s.seg_masks, s.camera_poses, s.mv_features, s.query_embeddings, return_obj=True, device="cuda")
        new.append((f.cpu(), w.cpu(), vis))
        ts.append((time.perf_counter() - t0) * 1e3)
    outs = new
    print("step", step, " ".join(f"{t:6.1f}" for t in ts))
t0 = time.perf_counter()
n = 0
for step in range(3):
    for (f, w, vis), _ in M.fuse_many([(s.points, s.colors, s.labels, s.depths, s.seg_masks, s.camera_poses, s.mv_features, s.query_embeddings) for s in host], return_obj=True, device="cuda"):
        n += 1
print("fuse_many ms/scene", (time.perf_counter() - t0) / n * 1e3)
