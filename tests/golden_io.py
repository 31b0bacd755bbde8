"""Loads tests/golden/*.npz back into the containers the fusion API takes."""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def intr_of(z):
    h, w, fx, fy, cx, cy = [float(x) for x in z["intr"]]
    return {"height": int(h), "width": int(w), "fx": fx, "fy": fy, "cx": cx, "cy": cy}


def scene_of(z, pixel=False):
    intr = intr_of(z)
    V = z["depths"].shape[0]
    if pixel:
        feats = [torch.from_numpy(f.copy()) for f in z["feats"]]
    else:
        rows = z["feat_rows"]
        off = np.r_[0, np.cumsum(rows)]
        allf = torch.from_numpy(z["feats"].copy())
        feats = [allf[off[v]:off[v + 1]].clone() for v in range(V)]
    return SimpleNamespace(
        points=z["points"].copy(), colors=z["colors"].astype(np.float64), labels=z["labels"].astype(np.int64),
        depths=[d.copy() for d in z["depths"]], seg_masks=[s.astype(np.int64) for s in z["segs"]],
        camera_poses=[p.copy() for p in z["poses"]], inv_poses=[p.copy() for p in z["inv_poses"]],
        mv_features=feats, query_embeddings=torch.from_numpy(z["query"].copy()), intrinsic=intr,
        n_views=V, n_points=z["points"].shape[0])


def unpack(bits, n):
    return np.unpackbits(bits, axis=-1)[..., :n]


def fp64_pose_case(seed=4321):
    """A scene with fp64 camera poses that are NOT fp32-representable plus points on depth-discontinuity pixel borders
    (projected u within ~1e-13 of an integer column whose two neighbours differ in depth): rounding the pose to fp32
    moves u by ~1e-5 px and flips about half of those points' visibility. Returns (scene, poses64, points)."""
    from dropclip_b200.scenes import small_scene
    sc = small_scene(seed, n_views=5, n_points=20000, n_objects=6, height=120, width=160)
    rng = np.random.default_rng(1)
    poses64, extra = [], []
    intr = sc.intrinsic
    for P, depth in zip(sc.camera_poses, sc.depths):
        P64 = P.astype(np.float64)
        P64[:3, 3] += rng.uniform(-1e-4, 1e-4, size=3)
        P64[:3, :3] += rng.uniform(-1e-9, 1e-9, size=(3, 3))
        poses64.append(P64)
        j, k = np.nonzero(np.abs(np.diff(depth.astype(np.float64), axis=1)) > 0.2)
        k = k + 1
        z = depth[j, k].astype(np.float64)
        xc = (k - intr["cx"]) / intr["fx"] * z
        yc = (j + 0.5 - intr["cy"]) / intr["fy"] * z
        cam = np.stack([xc, -yc, -z, np.ones_like(z)])  # Blender camera frame (utils/feature_fusion.py:75-79 flips back)
        extra.append((P64 @ cam)[:3].T)
    return sc, poses64, np.concatenate([sc.points] + extra)
