"""Where the host time of the upload path goes. One script, three views of the same question:

    python benchmarks/upload_probe.py groups    # upload_list of cold scenes: helper calls vs H2D enqueue vs drain, per group size
    python benchmarks/upload_probe.py calls     # upload_list inside fuse() vs called alone on the same arrays
    python benchmarks/upload_probe.py helpers   # the staging helper calls themselves inside fuse(), alone, and alone with pauses

(`helpers` is the run quoted in DESIGN.md SS5: helper calls of one scene take 7.9 ms behind a torch CPU copy, 3.3 ms without.)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dropclip_b200 import _lib
from dropclip_b200.engine import PinnedStaging

H, W = 480, 640


def production_fusion():
    from dropclip_b200.feature_fusion import MultiviewFeatureFusion
    from dropclip_b200.scenes import make_scene
    scs = [make_scene(1234 + i, n_views=73, n_points=100_000, n_objects=21, device="cuda") for i in range(3)]
    M = MultiviewFeatureFusion(scs[0].intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1,
                               norm_feat=False, device="cuda")

    def fuse(sc):
        return M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                      sc.query_embeddings, return_obj=True, device="cuda")
    return scs, fuse


def upload_alone(st, sc, pause=0.0):
    st.begin()
    st.upload_list(sc.depths, torch.float32, (H, W))
    if pause:
        time.sleep(pause)
    st.upload_list(sc.seg_masks, torch.int64, (H, W), narrow_to_u8=True)
    st.end()


def mode_groups():
    rng = np.random.default_rng(0)
    scenes = [([rng.random((H, W), dtype=np.float32) for _ in range(73)],
               [rng.integers(0, 21, size=(H, W), dtype=np.int64) for _ in range(73)]) for _ in range(4)]
    st = PinnedStaging("cuda")
    for gmb in (8, 32, 128, 1024):
        tot = []
        for _ in range(2):
            for depths, segs in scenes:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                st.begin()
                st.upload_list(depths, torch.float32, (H, W), group_bytes=gmb << 20)
                t1 = time.perf_counter()
                st.upload_list(segs, torch.int64, (H, W), group_bytes=gmb << 20, narrow_to_u8=True)
                t2 = time.perf_counter()
                st.end()
                torch.cuda.synchronize()
                tot.append((t1 - t0, t2 - t1, time.perf_counter() - t2))
        a = np.array(tot) * 1e3
        print(f"group {gmb:5d} MB: depth host {a[:, 0].mean():5.2f} ms | seg host {a[:, 1].mean():5.2f} ms | "
              f"drain {a[:, 2].mean():5.2f} ms | total {a.sum(1).mean():5.2f} ms")


def mode_calls():
    log = []
    orig = PinnedStaging.upload_list

    def timed(self, arrays, *a, **k):
        t0 = time.perf_counter()
        r = orig(self, arrays, *a, **k)
        log.append((len(arrays), str(arrays[0].dtype), round((time.perf_counter() - t0) * 1e3, 2)))
        return r
    PinnedStaging.upload_list = timed
    scs, fuse = production_fusion()
    for _ in range(3):
        for sc in scs:
            t0 = time.perf_counter()
            fuse(sc)
            dt = (time.perf_counter() - t0) * 1e3
        print("fuse ms", round(dt, 2), "upload_list calls:", log[-2:])
    print("alignment of sources:", scs[0].depths[0].ctypes.data % 64, scs[0].seg_masks[0].ctypes.data % 64)
    st = PinnedStaging("cuda")
    for _ in range(3):
        for sc in scs:
            torch.cuda.synchronize()
            upload_alone(st, sc)
        print("alone:", log[-2:])


def mode_helpers():
    lib = _lib.load()
    calls = []
    for name in ("dc_host_gather_copy", "dc_host_gather_narrow_i64_u8"):
        def wrapped(*a, _f=getattr(lib, name), _n=name[8:]):
            t0 = time.perf_counter()
            r = _f(*a)
            calls.append((_n, (time.perf_counter() - t0) * 1e3))
            return r
        setattr(lib, name, wrapped)

    def summary():
        c = [ms for n, ms in calls if n.startswith("gather_copy")]
        w = [ms for n, ms in calls if n.startswith("gather_narrow")]
        return (f"copy calls {len(c)} sum {sum(c):.2f} ms max {max(c):.2f} | narrow calls {len(w)} sum {sum(w):.2f} ms "
                f"max {max(w):.2f}")
    scs, fuse = production_fusion()
    for _ in range(3):
        for sc in scs:
            calls.clear()
            fuse(sc)
        print("in fuse :", summary())
    st = PinnedStaging("cuda")
    for _ in range(2):
        for sc in scs:
            torch.cuda.synchronize()
            calls.clear()
            upload_alone(st, sc)
        print("alone   :", summary())
    # alone, but with the pauses fuse() has between the calls (kernels + read-back): do sleeping workers wake up slowly?
    for _ in range(2):
        for sc in scs:
            torch.cuda.synchronize()
            calls.clear()
            time.sleep(0.004)
            upload_alone(st, sc, pause=0.001)
        print("w/ pause:", summary())


if __name__ == "__main__":
    {"groups": mode_groups, "calls": mode_calls, "helpers": mode_helpers}[sys.argv[1] if len(sys.argv) > 1 else "groups"]()
