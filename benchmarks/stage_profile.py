"""cProfile of the host-side staging of one MV-TOD-sized scene (SceneBatch.from_host)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200.feature_fusion import MultiviewFeatureFusion
from dropclip_b200.engine import SceneBatch, PinnedStaging, FusionEngine

sc = make_scene(1234, n_views=73, n_points=100_000, n_objects=21, device="cuda")
M = MultiviewFeatureFusion(sc.intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
args = (sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings)
for _ in range(3):
    M.fuse(*args, return_obj=True, device="cuda")
pr = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.perf_counter()
pr.enable()
for _ in range(5):
    M.fuse(*args, return_obj=True, device="cuda")
pr.disable()
torch.cuda.synchronize()
print("ms/scene", (time.perf_counter() - t0) / 5 * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
