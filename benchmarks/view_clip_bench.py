"""generate_view_clip on the GPU: N = 100 k points x V views x 768-d, 24x32 patch map -> (V, N, 768) fp32.
Algorithmic bytes = the output rows (N * C * 4 per view; the 2.4 MB patch map stays in L2)."""

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dropclip_b200 import sample_builder as sb


def main():
    rng = np.random.default_rng(0)
    n, c = 100_000, 768
    pc = rng.uniform(-4, 4, size=(n, 3)) * [1, 1, 0.2]
    K = np.array([[444.44444444, 0, 319.5], [0, 444.44444444, 239.5], [0, 0, 1]])
    for v in (1, 8):
        poses = []
        for i in range(v):
            a = 2 * np.pi * i / max(v, 1)
            m = np.eye(4)
            m[:3, 3] = [0.0, 0.0, 14.0]
            m[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
            poses.append(m)
        patch = torch.randn(v, 24, 32, c, device="cuda")
        d_pc = torch.from_numpy(pc).cuda()
        for _ in range(3):
            out = sb.generate_view_clips(d_pc, np.stack(poses), K, patch, return_device=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            out = sb.generate_view_clips(d_pc, np.stack(poses), K, patch, return_device=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gb = v * n * c * 4 / 1e9
        print(f"V={v}: {ms:.3f} ms per call (incl. host-side pose inversion/upload), {gb / ms * 1e3:.0f} GB/s written, "
              f"{ms / v * 1e3:.0f} us per view")


if __name__ == "__main__":
    main()
