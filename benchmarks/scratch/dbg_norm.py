import os, sys
sys.path.insert(0, "/root/repo")
import torch
from dropclip_b200 import _lib
from dropclip_b200.engine import FusionEngine
eng = FusionEngine("cuda")
g = torch.Generator(device="cuda").manual_seed(77)
n, c = 120_000, 768
x = (torch.randn((n, c), generator=g, device="cuda") * torch.rand((n, 1), generator=g, device="cuda") * 4).half()
x[5] = 0; x[6] = 6e-8; x[7] = 2000.0; x[8] = 0; x[8, 3] = 1.0009765625
d = x.double().pow(2).sum(-1, keepdim=True).sqrt()
h0 = d.float().half(); bits = h0.view(torch.int16)
cands = torch.stack([h0, (bits + 1).view(torch.float16), (bits - 1).clamp_min(0).view(torch.float16)], 0)
pick = (cands.double() - d).abs().nan_to_num(nan=float("inf")).argmin(0, keepdim=True)
want_nrm = torch.where(torch.isfinite(h0) & (h0 > 0), cands.gather(0, pick)[0], h0)
want = (x.float() / want_nrm.float()).half()
y = x.clone(); t = torch.randn((4, c), generator=g, device="cuda").half()
eng.ground(y, t, _lib.DC_GROUND_RAW, 0.1, normalize=True); torch.cuda.synchronize()
same = (y.view(torch.int16) == want.view(torch.int16)) | (torch.isnan(y) & torch.isnan(want))
bad = (~same).any(-1).nonzero().view(-1)
print("bad rows", bad.tolist())
for r in bad.tolist()[:6]:
    cols = (~same[r]).nonzero().view(-1)
    print(r, "ncols", cols.numel(), "d", d[r].item(), "want_nrm", want_nrm[r].item(), "h0", h0[r].item())
    k = cols[0].item()
    print("  x", x[r, k].item(), "y", y[r, k].item(), "want", want[r, k].item(), "implied nrm", (x[r,k].float()/y[r,k].float()).item())
