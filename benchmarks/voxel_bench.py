"""(5) voxelisation at the training shape: B samples x 10 000 points (data/dataset_blender.py:20 MAX_POINTS),
features 768 + 3 + 3, voxel size 0.05 (config/DistilBlender.yaml:6). Times quantise and the feature gather."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine

B, N, F = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 10_000, 774
eng = FusionEngine("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
xyz = (torch.rand((B * N, 3), generator=g, device="cuda") - 0.5) * 4.0
feats = torch.randn((B * N, F), generator=g, device="cuda")
labels = torch.randint(0, 21, (B * N,), generator=g, device="cuda", dtype=torch.int32)
off = torch.arange(0, B * N + 1, N, dtype=torch.int64, device="cuda")
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
def ev(): return torch.cuda.Event(enable_timing=True)
for vs in (0.05, 0.02):
    tq, tg = [], []
    for it in range(6):
        a, b, c = ev(), ev(), ev()
        a.record()
        vox = eng.voxelize(xyz, off, vs, labels, 0)
        b.record()
        m = int(vox["voxel_off"][-1].item())
        b2 = ev(); b2.record()
        out = eng.voxel_gather(feats, off, vox, m)
        c.record(); torch.cuda.synchronize()
        tq.append(a.elapsed_time(b)); tg.append(b2.elapsed_time(c))
    q, gth = sorted(tq[1:])[2], sorted(tg[1:])[2]
    qbytes = B * N * (12 + 4 + 8) + m * (12 + 8 + 4)
    gbytes = 2 * m * F * 4 + m * 8
    print(json.dumps({"case": f"voxelize_B{B}x10k_vs{vs}", "voxels": m, "quantize_ms": q, "quantize_alg_GBps": qbytes / q / 1e6,
                      "gather_ms": gth, "gather_GBps": gbytes / gth / 1e6, "gather_frac_hbm": gbytes / gth / 1e6 / peaks.get("hbm_gbs", 6650.0),
                      "samples_per_s": B / ((q + gth) * 1e-3)}))
