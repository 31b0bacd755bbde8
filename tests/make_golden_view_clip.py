"""Writes tests/golden/view_clip.npz: seeded inputs and the outputs of the UNMODIFIED reference method
`MVDistilDataset.generate_view_clip` (data/dataset_blender.py:132-171), called unbound on a stand-in `self` that
carries only what the method reads (root with a cameras json, K, patch_h/patch_w and a stub CLIP extractor that
returns a seeded (patch_h * patch_w, C) tensor). Run in the build container only
(`python tests/make_golden_view_clip.py`); the reference tree does not travel to the GPU box."""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def look_at(eye, target=(0.0, 0.0, 0.0)):
    """camera->world matrix, Blender convention (camera looks along -Z, +Y up)."""
    eye, target = np.asarray(eye, dtype=np.float64), np.asarray(target, dtype=np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
    return m


def cases():
    """(name, pc, world_matrix, K, patch (ph, pw, C), h, w)"""
    out = []
    rng = np.random.default_rng(21)
    K = np.array([[444.44444444, 0, 319.5], [0, 444.44444444, 239.5], [0, 0, 1]])
    pc = rng.uniform(-4, 4, size=(400, 3)) * [1, 1, 0.25]
    out.append(("scene", pc, look_at((9.0, -7.0, 8.0)), K, rng.standard_normal((24, 32, 16)).astype(np.float32), 480, 640))
    # points behind the camera, far outside the image (clipped), on the camera plane (z' == 0 -> pixel (0,0)),
    # non-finite coordinates (-> INT64_MIN -> clipped to 0)
    wild = rng.uniform(-30, 30, size=(300, 3))
    wild[:10] = [[1.5, -2.0, 0.0]] * 10
    wild[:10, :2] += rng.uniform(-1, 1, size=(10, 2))
    wild[10] = [np.nan, 0.0, 1.0]
    wild[11] = [np.inf, 0.0, 1.0]
    wild[12] = [0.0, -np.inf, 2.0]
    wild[13] = [1e300, 1e300, -1e300]
    out.append(("wild", wild, np.eye(4), K, rng.standard_normal((24, 32, 12)).astype(np.float32), 480, 640))
    Ks = np.array([[70.0, 0, 39.5], [0, 70.0, 29.5], [0, 0, 1]])
    pc2 = rng.uniform(-3, 3, size=(257, 3))
    out.append(("small_odd_dim", pc2, look_at((5.0, 5.0, 6.0)), Ks, rng.standard_normal((6, 8, 18)).astype(np.float32), 60, 80))
    out.append(("downsample", pc2, look_at((-6.0, 2.0, 5.0)), Ks, rng.standard_normal((90, 100, 8)).astype(np.float32), 60, 80))
    return out


def main():
    ref_shim.load()
    for m in ("h5py", "MinkowskiEngine"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import data.dataset_blender as db
    g = {}
    with tempfile.TemporaryDirectory() as root:
        for name, pc, wm, K, patch, h, w in cases():
            os.makedirs(os.path.join(root, name), exist_ok=True)
            with open(os.path.join(root, name, f"cameras.{name}.json"), "w") as f:
                json.dump({"view003": {"world_matrix": wm.tolist()}}, f)
            ph, pw, c = patch.shape
            stub = types.SimpleNamespace(root=root, K=K, patch_h=ph, patch_w=pw,
                                         CLIP=types.SimpleNamespace(extract=lambda files, p=patch: [torch.from_numpy(p).reshape(-1, p.shape[-1])]))
            with np.errstate(all="ignore"):
                feats = db.MVDistilDataset.generate_view_clip(stub, pc.copy(), name, 3, h=h, w=w)
            assert feats.shape == (pc.shape[0], c) and feats.dtype == torch.float32
            g[f"{name}_pc"], g[f"{name}_world_matrix"], g[f"{name}_K"] = pc, wm, K
            g[f"{name}_patch"], g[f"{name}_hw"], g[f"{name}_out"] = patch, np.array([h, w]), feats.numpy()
    np.savez_compressed(os.path.join(OUT, "view_clip.npz"), **g)
    print("wrote", os.path.join(OUT, "view_clip.npz"), os.path.getsize(os.path.join(OUT, "view_clip.npz")), "bytes")


if __name__ == "__main__":
    main()
