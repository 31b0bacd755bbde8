"""One pixel-level fusion call (V views, N=100k, 24x32x768 maps, sim max, norm_feat) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine, batch_from_device
from dropclip_b200.scenes import make_scene
V = int(sys.argv[1]) if len(sys.argv) > 1 else 73
eng = FusionEngine("cuda")
sc = make_scene(1234, n_views=V, n_points=100_000, n_objects=21, device="cuda", as_torch=True, pixel_features=True, feature_dtype=torch.float32)
objs = sc["mv_features"]
sc_obj = dict(sc)
sc_obj["mv_features"] = [torch.zeros((1, 768), device="cuda", dtype=torch.float16) for _ in range(V)]
b = batch_from_device([sc_obj], "cuda")
b.feats = torch.stack(objs).contiguous()
mask, any_vis, _ = eng.visibility(b, 0.05, torch.uint8)
for _ in range(int(os.environ.get("REPS", "2"))):
    sums, w = eng.pixel_fuse(b, mask, "max", True, normalize=True)
torch.cuda.synchronize()
print("ok")
