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


# ------------------------------------------------------------------------------------------------
# REGRAD-style helpers (utils/projections.py:151-241, utils/geometry.py:350-352,390-401).
# PARITY UNPINNED for voxel_down_ref: Open3D 0.15.2 is not installable here and the reference has
# no fixture; the function restates Open3D's published voxel_down_sample semantics (voxel index =
# floor((p - (min_bound - size/2)) / size), output = mean of the voxel's points summed in point
# order). Output is ordered by voxel index (Open3D's own order is hash-map iteration order).
def voxel_down_ref(points: np.ndarray, voxel_size: float):
    pts = np.asarray(points, dtype=np.float64)
    origin = pts.min(axis=0) - voxel_size * 0.5
    idx = np.floor((pts - origin) / voxel_size).astype(np.int64)
    key = (idx[:, 0] << 42) | (idx[:, 1] << 21) | idx[:, 2]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    heads = np.r_[True, ks[1:] != ks[:-1]]
    starts = np.flatnonzero(heads)
    ends = np.r_[starts[1:], len(ks)]
    out = np.empty((len(starts), 3))
    first = np.empty(len(starts), dtype=np.int64)
    for u, (a, b) in enumerate(zip(starts, ends)):
        s = np.zeros(3)
        for j in order[a:b]:  # sequential accumulation in point order, like Open3D's AccumulatedPoint
            s = s + pts[j]
        out[u] = s / float(b - a)
        first[u] = order[a]
    return out, first


def nearest_ref(full_pc, filtered_pc):
    import scipy.spatial
    return scipy.spatial.cKDTree(full_pc).query(filtered_pc)[1]


def camera_points_ref(pc, pose):
    inv = np.linalg.inv(pose)
    return np.dot(inv, np.vstack([pc.T, np.ones((1, pc.shape[0]))]))[:3, :].T


def fuse_multiview_ref(pcs, feats, poses, intr, crop_size=336, patch_size=14, voxel_size=0.0075, reshape_feat=False,
                       norm_feat=True):
    """utils/projections.py:151-211 with pc_voxel_down -> voxel_down_ref."""
    pc_aggr, _ = voxel_down_ref(np.concatenate(pcs, axis=0), voxel_size)
    n, C = pc_aggr.shape[0], feats.shape[-1]
    ph = pw = crop_size // patch_size
    sums = torch.zeros((n, C), dtype=float)
    counter = torch.zeros((n, 1), dtype=float)
    for pc, feat, pose in zip(pcs, feats, poses):
        ids, first = np.unique(nearest_ref(pc_aggr, pc), return_index=True)
        cam = camera_points_ref(pc, pose)
        cam[:, 2] = -cam[:, 2]
        cam[:, 1] = -cam[:, 1]
        mapping = pixels_of(cam, intr)
        pixels = mapping[first].squeeze().astype(int)
        if len(pixels.shape) < 2:
            continue
        ys = np.clip(pixels[:, 1], 0, intr["height"] - 1)
        xs = np.clip(pixels[:, 0], 0, intr["width"] - 1)
        if reshape_feat:
            feat = feat.reshape(ph, pw, C)
        if norm_feat:
            feat /= feat.norm(dim=-1, keepdim=True)
        full = nearest_patch_map(feat, (intr["height"], intr["width"], 3))
        sums[ids, :] = sums[ids, :] + full[ys, xs]
        counter[ids, :] += 1
    counter[counter == 0] = 1e-5
    return sums / counter, pc_aggr


def rgbd_points_ref(rgb, depth, intr, depth_scale=1.0, depth_trunc=25.0):
    """Open3D create_from_rgbd_image (convert_rgb_to_intensity=False) restated. PARITY UNPINNED."""
    d = depth.astype(np.float32) / np.float32(depth_scale)
    d = np.where(d >= depth_trunc, np.float32(0), d)
    v, u = np.nonzero(d > 0)
    z = d[v, u].astype(np.float64)
    x = (u - intr["cx"]) * z / intr["fx"]
    y = (v - intr["cy"]) * z / intr["fy"]
    return np.stack([x, y, z], axis=1), rgb[v, u].astype(np.float64) / 255.0


# ------------------------------------------------------------------------------------------------
# Aggregation step upstream of the fusion path: aggregate_views_blender_new
# utils/geometry.py:120-204. PARITY UNPINNED (Open3D create_from_rgbd_image / transform /
# voxel_down_sample_and_trace are restated from their published semantics; no fixture exists).
def voxel_down_trace_ref(points, colors, labels, voxel_size):
    """Per voxel: mean point, mean colour, Counter(labels).most_common()[0][0], member count; voxel order = key order."""
    from collections import Counter
    pts = np.asarray(points, dtype=np.float64)
    origin = pts.min(axis=0) - voxel_size * 0.5
    idx = np.floor((pts - origin) / voxel_size).astype(np.int64)
    key = (idx[:, 0] << 42) | (idx[:, 1] << 21) | idx[:, 2]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    ends = np.r_[starts[1:], len(ks)]
    out_p = np.empty((len(starts), 3))
    out_c = np.empty((len(starts), 3))
    out_l = np.empty(len(starts), dtype=np.int64)
    cnt = np.empty(len(starts), dtype=np.int64)
    for u, (a, b) in enumerate(zip(starts, ends)):
        sp, sc = np.zeros(3), np.zeros(3)
        for j in order[a:b]:
            sp = sp + pts[j]
            sc = sc + colors[j]
        out_p[u], out_c[u] = sp / float(b - a), sc / float(b - a)
        out_l[u] = Counter(labels[order[a:b]].tolist()).most_common()[0][0]  # utils/geometry.py:197
        cnt[u] = b - a
    return out_p, out_c, out_l, cnt


def binary_masks_to_seg_ref(masks, obj_ids):
    """utils/image.py:11-15."""
    return np.max(masks * obj_ids[:, None, None], axis=0)


def aggregate_views_ref(scene, intr, depth_trunc=25.0, voxel_size=None):
    """utils/geometry.py:120-204 with the Open3D calls restated."""
    pts, cols, labs = [], [], []
    for _, stuff in scene["views"].items():
        rgb, depth = stuff["rgb"], stuff["depth"].astype(np.float32)
        valid = depth < depth_trunc
        _, masks, colors = zip(*stuff["annos"])
        seg = binary_masks_to_seg_ref(np.stack(masks), np.asarray([scene["col_to_ins"][c] for c in colors]))
        labs.append(seg[valid].flatten())
        p, c = rgbd_points_ref(rgb, depth, intr, depth_trunc=depth_trunc)
        p = p * np.array([1.0, -1.0, -1.0])  # T_cam
        m = np.asarray(stuff["camera"]["world_matrix"]).astype(np.float64)
        p = p @ m[:3, :3].T + m[:3, 3]
        pts.append(p)
        cols.append(c)
    pts, cols, labs = np.concatenate(pts), np.concatenate(cols), np.concatenate(labs)
    if voxel_size is None:
        return pts, cols, labs
    p, c, l, _ = voxel_down_trace_ref(pts, cols, labs, voxel_size)
    return p, c, l
