"""BASELINE config 4 (pixel-level vs object-level ablation): one MV-TOD-shaped scene through the pixel path
(V views, 24x32x768 patch maps, N=100k points) + voxel-size sweep of the voxeliser."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine, batch_from_device
from dropclip_b200.scenes import make_scene

def ev():
    return torch.cuda.Event(enable_timing=True)

def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    eng = FusionEngine("cuda")
    sc = make_scene(1234, n_views=V, n_points=100_000, n_objects=21, device="cuda", as_torch=True, pixel_features=True,
                    feature_dtype=torch.float32)
    objs = sc["mv_features"]
    sc_obj = dict(sc)
    sc_obj["mv_features"] = [torch.zeros((1, 768), device="cuda", dtype=torch.float16) for _ in range(V)]
    b = batch_from_device([sc_obj], "cuda")
    b.feats = torch.stack(objs).contiguous()
    mask, any_vis, _ = eng.visibility(b, 0.05, torch.uint8)
    for kern, nf in (("max", True), (None, True), ("max", False)):
        ts = []
        for it in range(4):
            a, e = ev(), ev()
            a.record()
            sums, w = eng.pixel_fuse(b, mask, kern, nf, normalize=True)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        nvis = int(mask.sum().item())
        ms = sorted(ts[1:])[len(ts[1:]) // 2]
        print(json.dumps({"case": f"pixel_fuse_V{V}_sim_{kern}_norm_{nf}", "ms": ms, "visible_point_views": nvis,
                          "scenes_per_s": 1e3 / ms, "gflops_taps": 2 * 16 * 768 * nvis / (ms * 1e-3) / 1e9}))
    # voxel-size sweep (scripts/RUN_voxel_abls.bash sizes x world_scale 10, and the training size 0.05)
    xyz = sc["points"].float().contiguous()
    off = torch.tensor([0, xyz.shape[0]], dtype=torch.int64, device="cuda")
    for vs in (0.02, 0.04, 0.06, 0.08, 0.05):
        ts = []
        for it in range(4):
            a, e = ev(), ev()
            a.record()
            vox = eng.voxelize(xyz, off, vs, sc["labels"].int(), 0)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        print(json.dumps({"case": f"voxelize_100k_vs{vs}", "ms": min(ts[1:]), "voxels": int(vox["voxel_off"][-1].item())}))

if __name__ == "__main__":
    main()
