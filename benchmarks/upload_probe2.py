"""upload_list timings inside fuse() vs called alone on the same arrays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200.scenes import make_scene
from dropclip_b200.feature_fusion import MultiviewFeatureFusion
from dropclip_b200.engine import PinnedStaging

log = []
orig = PinnedStaging.upload_list
def timed(self, arrays, *a, **k):
    t0 = time.perf_counter()
    r = orig(self, arrays, *a, **k)
    log.append((len(arrays), str(arrays[0].dtype), (time.perf_counter() - t0) * 1e3))
    return r
PinnedStaging.upload_list = timed

scs = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda") for i in range(3)]
M = MultiviewFeatureFusion(scs[0].intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
for rep in range(3):
    for sc in scs:
        t0 = time.perf_counter()
        M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings, return_obj=True, device="cuda")
        dt = (time.perf_counter() - t0) * 1e3
    print("fuse ms", round(dt, 2), "upload_list calls:", [(n, d, round(ms, 2)) for n, d, ms in log[-2:]])
print("alignment of sources:", scs[0].depths[0].ctypes.data % 64, scs[0].seg_masks[0].ctypes.data % 64, type(scs[0].depths), scs[0].depths[0].flags.owndata)
st = PinnedStaging("cuda")
for rep in range(3):
    for sc in scs:
        torch.cuda.synchronize()
        st.begin()
        d = st.upload_list(sc.depths, torch.float32, (480, 640))
        s, ok = st.upload_list(sc.seg_masks, torch.int64, (480, 640), narrow_to_u8=True)
        st.end()
    print("alone:", [(n, dd, round(ms, 2)) for n, dd, ms in log[-2:]])
