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
