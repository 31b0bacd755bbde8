"""Where does the end-to-end time of MultiviewFeatureFusion.fuse() go? (host staging, H2D, kernels, D2H)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200.feature_fusion import MultiviewFeatureFusion
from dropclip_b200.engine import SceneBatch, PinnedStaging, FusionEngine

def T():
    torch.cuda.synchronize(); return time.perf_counter()

sc = make_scene(1234, n_views=73, n_points=100_000, n_objects=21, device="cuda")
M = MultiviewFeatureFusion(sc.intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
args = (sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings)
for _ in range(3): M.fuse(*args, return_obj=True, device="cuda")
t0 = T()
for _ in range(5): M.fuse(*args, return_obj=True, device="cuda")
print("fuse total ms/scene", (T() - t0) / 5 * 1e3)

eng, st = FusionEngine("cuda"), PinnedStaging("cuda")
scene = M._scene(sc.points, sc.depths, sc.camera_poses, sc.labels, sc.seg_masks, sc.mv_features, sc.query_embeddings)
for rep in range(3):
    t0 = T(); b = SceneBatch.from_host([scene], "cuda", staging=st); t1 = T()
    res = eng.fuse_object_level(b, 0.05, False, True, "max", torch.uint8); t2 = T()
    comp = eng.compact_visibility(b, res["any_visible"], res["records"], res["rank"], torch.int64); t3 = T()
    keep = res["any_visible"].cpu().numpy().astype(bool); p = sc.points[keep]; c = sc.colors[keep]; l = sc.labels[keep]; t4 = T()
    m64 = torch.empty((73, int(comp[2][-1])), dtype=torch.int64, pin_memory=True); t5 = T()
    m64.copy_(comp[4].view(73, -1)); t6 = T()
    f = res["fused"].cpu(); w = res["weight_obj"].cpu(); t7 = T()
    print(f"from_host {1e3*(t1-t0):.2f} | kernels {1e3*(t2-t1):.2f} | compact {1e3*(t3-t2):.2f} | host filter {1e3*(t4-t3):.2f} | "
          f"pinned alloc {1e3*(t5-t4):.2f} | i64 mask D2H {1e3*(t6-t5):.2f} | feats D2H {1e3*(t7-t6):.2f}")
# raw copy speeds
big = torch.empty(275_000_000, dtype=torch.uint8, pin_memory=True); dev = torch.empty_like(big, device="cuda")
t0 = T(); dev.copy_(big, non_blocking=True); t1 = T(); print("H2D pinned 275MB ms", 1e3*(t1-t0), "GB/s", 0.275/(t1-t0))
src = torch.from_numpy(np.stack(sc.seg_masks))
t0 = T(); big[:src.numel()*8].view(torch.int64).view(src.shape).copy_(src); t1 = T(); print("host->pinned 180MB one copy_ ms", 1e3*(t1-t0), "GB/s", 0.18/(t1-t0))
t0 = T()
for i, m in enumerate(sc.seg_masks):
    big[i*2457600:(i+1)*2457600].view(torch.int64).view(480, 640).copy_(torch.from_numpy(m))
t1 = T(); print("host->pinned 73 copies ms", 1e3*(t1-t0))
from concurrent.futures import ThreadPoolExecutor
def cp(i): big[i*2457600:(i+1)*2457600].view(torch.int64).view(480, 640).copy_(torch.from_numpy(sc.seg_masks[i]))
for nt in (2, 4, 8):
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(cp, range(73)))
        t0 = T(); list(ex.map(cp, range(73))); t1 = T(); print(f"host->pinned 73 copies, {nt} threads ms", 1e3*(t1-t0))
print("torch threads", torch.get_num_threads(), "cpus", os.cpu_count())
