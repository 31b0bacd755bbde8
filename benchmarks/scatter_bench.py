"""scatter_to_points (reconstruct_per_obj_feat / feat[label]): HBM-write-bound, N' * C * 4 bytes per scene."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200 import _lib
from dropclip_b200.engine import FusionEngine
from dropclip_b200._lib import check, ptr, current_stream

S, N, Q, C = 16, 100_000, 21, 768
eng = FusionEngine("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
fused = torch.randn((S * Q, C), generator=g, device="cuda")
labels = torch.randint(0, Q, (S * N,), generator=g, device="cuda")
q_off = torch.arange(0, S * Q + 1, Q, dtype=torch.int64, device="cuda")
p_off = torch.arange(0, S * N + 1, N, dtype=torch.int64, device="cuda")
out = torch.empty((S * N, C), dtype=torch.float32, device="cuda")
def run():
    check(eng.lib.dc_scatter_to_points(ptr(fused), ptr(q_off), ptr(labels), ptr(p_off), S, N, C, 1, ptr(out), current_stream()))
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
want = fused.view(S, Q, C)[torch.arange(S, device="cuda").repeat_interleave(N), labels] * (labels > 0).unsqueeze(1)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
bytes_ = S * N * (C * 4 + 8)
print(json.dumps({"case": "scatter_to_points", "scenes": S, "ms": ms, "GBps": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / peaks.get("hbm_gbs", 6650.0),
                  "equal": bool(torch.equal(out, want))}))
