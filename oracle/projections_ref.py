"""ORACLE (test infrastructure, never the product path): CPU restatement of the geometry
helpers of `utils/projections.py` / `utils/transforms.py` and of the MinkowskiEngine
voxelisation call the dataset makes.

Parity status:
  * back_project, pixels_of, to_world, unique_max_pool, nearest_patch_map: PINNED against the
    unmodified reference functions by `tests/make_golden.py` -> `tests/golden/proj_*.npz`.
  * sparse_quantize_ref: **PARITY UNPINNED**. MinkowskiEngine is an un-vendored, un-pinned
    third-party dependency (README.md:25; not in requirements.txt) that is not installable here
    and the reference has no test or fixture at that boundary. The function restates the
    published ME 0.5.x semantics (floor(xyz / size) in the input dtype -> int32, hash-unique with
    first-occurrence index, label collision -> ignore_label) and parity is anchored on the
    reference's call sites (data/dataset_blender.py:406-414, data/dataset.py:164-172) through
    order-insensitive properties.

Reference lines followed (under /root/reference):
  back_project        utils/projections.py:67-86
  pixels_of           utils/projections.py:59-64
  to_world            utils/transforms.py:43-49
  unique_max_pool     utils/projections.py:245-261
  nearest_patch_map   utils/transforms.py:149-165
"""
from __future__ import annotations

import numpy as np
import torch


def back_project(depth: np.ndarray, intr) -> np.ndarray:
    h, w = depth.shape
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    x = (u - intr["cx"]) / intr["fx"]
    y = (v - intr["cy"]) / intr["fy"]
    z = depth.copy()
    return np.stack((np.multiply(x, z), np.multiply(y, z), z), axis=-1)


def pixels_of(cam_points: np.ndarray, intr) -> np.ndarray:
    px = np.zeros((cam_points.shape[0], 2))
    px[:, 0] = intr["fx"] * cam_points[:, 0] / cam_points[:, 2] + intr["cx"]
    px[:, 1] = intr["fy"] * cam_points[:, 1] / cam_points[:, 2] + intr["cy"]
    return px


def to_world(cam_points: np.ndarray, pose: np.ndarray) -> np.ndarray:
    homog = np.vstack([cam_points.T, np.ones((1, cam_points.shape[0]))])
    return np.dot(pose, homog)[:3, :].T


def unique_max_pool(points: np.ndarray, feats: np.ndarray):
    uniq, inv = np.unique(points, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    ordered = feats[inv.argsort()]
    starts = np.r_[0, np.cumsum(np.bincount(inv))]
    return uniq, np.maximum.reduceat(ordered, starts[:-1], axis=0)


def nearest_patch_map(feat: torch.Tensor, image_shape):
    H, W, _ = image_shape
    ph, pw, _ = feat.shape
    y = torch.arange(H).unsqueeze(1).expand(H, W).float()
    x = torch.arange(W).unsqueeze(0).expand(H, W).float()
    return feat[(y * (ph / H)).long(), (x * (pw / W)).long()]


def sparse_quantize_ref(xyz: np.ndarray, features=None, labels=None, ignore_label: int = -100,
                        quantization_size=None):
    """ME.utils.sparse_quantize(..., return_index=True, return_inverse=True) restated.

    Canonical voxel order = order of first occurrence. Returns
    (coords int32 (M,3), features[unique_map] | None, voxel_labels | None, unique_map, inverse_map).
    """
    xyz = np.asarray(xyz)
    if quantization_size is not None:
        q = np.floor(xyz / np.asarray(quantization_size, dtype=xyz.dtype))
    elif np.issubdtype(xyz.dtype, np.floating):
        q = np.floor(xyz)
    else:
        q = xyz
    q = q.astype(np.int32)
    _, first, inv = np.unique(q, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")  # sorted-unique slot -> first-occurrence rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    unique_map = first[order].astype(np.int64)
    inverse_map = rank[inv].astype(np.int64)
    coords = q[unique_map]
    vox_labels = None
    if labels is not None:
        labels = np.asarray(labels)
        vox_labels = labels[unique_map].copy()
        clash = labels != vox_labels[inverse_map]
        vox_labels[np.unique(inverse_map[clash])] = ignore_label
    feats = None if features is None else np.asarray(features)[unique_map]
    return coords, feats, vox_labels, unique_map, inverse_map


def sparse_collate_ref(coords_list, feats_list):
    """ME.utils.sparse_collate: prepend an int32 batch-index column and concatenate."""
    out = []
    for b, c in enumerate(coords_list):
        c = np.asarray(c).astype(np.int32)
        out.append(np.concatenate([np.full((c.shape[0], 1), b, dtype=np.int32), c], axis=1))
    return np.concatenate(out, axis=0), np.concatenate([np.asarray(f) for f in feats_list], axis=0)
