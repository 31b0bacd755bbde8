"""One predict()-shaped grounding call (200k x 256 x 768 fp16, paired) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dropclip_b200 import _lib
from dropclip_b200.engine import FusionEngine
n, p, c = 200_000, 256, 768
eng = FusionEngine("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn((n, c), generator=g, device="cuda").half()
t = torch.randn((p, c), generator=g, device="cuda")
t = (t / t.norm(dim=-1, keepdim=True)).half()
for _ in range(int(os.environ.get("REPS", "2"))):
    x = x0.clone()
    out, pred = eng.predict(x, t, _lib.DC_GROUND_PAIRED, 0.1, True, 0.7)
torch.cuda.synchronize()
print("ok")
