"""Time of the staging helper calls themselves inside fuse() vs alone (same arrays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dropclip_b200 import _lib
from dropclip_b200.scenes import make_scene
from dropclip_b200.feature_fusion import MultiviewFeatureFusion
from dropclip_b200.engine import PinnedStaging

lib = _lib.load()
calls = []
def wrap(name):
    f = getattr(lib, name)
    def g(*a):
        t0 = time.perf_counter()
        r = f(*a)
        calls.append((name[8:], (time.perf_counter() - t0) * 1e3))
        return r
    setattr(lib, name, g)
wrap("dc_host_gather_copy"); wrap("dc_host_gather_narrow_i64_u8")

scs = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda") for i in range(3)]
M = MultiviewFeatureFusion(scs[0].intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, device="cuda")
def summary():
    c = [ms for n, ms in calls if n.startswith("gather_copy")]; w = [ms for n, ms in calls if n.startswith("gather_narrow")]
    return f"copy calls {len(c)} sum {sum(c):.2f} ms max {max(c):.2f} | narrow calls {len(w)} sum {sum(w):.2f} ms max {max(w):.2f}"
for rep in range(3):
    for sc in scs:
        calls.clear()
        M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings, return_obj=True, device="cuda")
    print("in fuse :", summary())
st = PinnedStaging("cuda")
for rep in range(2):
    for sc in scs:
        torch.cuda.synchronize(); calls.clear()
        st.begin()
        d = st.upload_list(sc.depths, torch.float32, (480, 640))
        s, ok = st.upload_list(sc.seg_masks, torch.int64, (480, 640), narrow_to_u8=True)
        st.end()
    print("alone   :", summary())
# alone, but with the pause fuse() has between calls (kernels + read-back ~2 ms): do sleeping worker threads wake up slowly?
for rep in range(2):
    for sc in scs:
        torch.cuda.synchronize(); calls.clear(); time.sleep(0.004)
        st.begin()
        d = st.upload_list(sc.depths, torch.float32, (480, 640))
        time.sleep(0.001)
        s, ok = st.upload_list(sc.seg_masks, torch.int64, (480, 640), narrow_to_u8=True)
        st.end()
    print("w/ pause:", summary())
