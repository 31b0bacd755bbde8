"""Instance-histogram kernels alone (headline batch): register-staged vs ring variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200.engine import FusionEngine, batch_from_device
from dropclip_b200.scenes import make_scene
dev = torch.device("cuda", 0)
eng = FusionEngine(dev)
uniq = [make_scene(1234 + i, n_views=73, n_points=1000, n_objects=21, device="cuda:0", as_torch=True) for i in range(16)]
batch = batch_from_device([uniq[i % 16] for i in range(64)], dev, seg_dtype=torch.int64)
gb = batch.segs.numel() * 8 / 1e9
configs = [("ldg", 0, 0, 8), ("ring", 6, 0, 8), ("ring", 6, 1, 8), ("ring", 4, 0, 8), ("ring", 9, 0, 8), ("ring", 6, 0, 12), ("ring", 4, 0, 12),
           ("ring", 8, 0, 16), ("ring", 12, 0, 16), ("ring", 8, 1, 16)]
if len(sys.argv) > 1:
    configs = [tuple(int(x) if x.isdigit() else x for x in a.split(",")) for a in sys.argv[1:]]
os.environ["DC_SEG_MODE"] = "ldg"
ref = [t.clone() for t in eng.seg_tables(batch)]
for mode, stages, flags, warps in configs:
    os.environ.update(DC_SEG_MODE=mode, DC_SEG_STAGES=str(stages), DC_SEG_FLAGS=str(flags), DC_SEG_WARPS=str(warps))
    for _ in range(3): got = eng.seg_tables(batch)
    torch.cuda.synchronize()
    ok = all(torch.equal(a, b) for a, b in zip(ref, got))
    eng.profile = {}
    for _ in range(10): eng.seg_tables(batch)
    torch.cuda.synchronize()
    ms = eng.profile_ms()["seg_histogram"]
    eng.profile = None
    print(f"{mode} depth={stages} flags={flags} warps={warps}: {ms:.3f} ms  {gb / ms:.2f} TB/s  tables {'ok' if ok else 'DIFFER'}", flush=True)
