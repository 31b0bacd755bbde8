"""Training-sample assembly on the GPU: the deterministic part of
`MVDistilDataset.__getitem__` / `collate_fn` (data/dataset_blender.py:330-362, 400-414, 437-461).

Per sample the reference gathers `feat = per_obj[label]` for the whole cloud on the CPU (307 MB at
100 k points), filters by the visibility of the chosen views, draws MAX_POINTS random points, centres
them, concatenates [feat, xyz, rgb] and voxelises with MinkowskiEngine. `build_samples` does the same
for a batch of samples in a handful of launches and gathers only the selected rows. The random choices
of the reference (`view_ids`, `indices`) are arguments, so a DataLoader keeps drawing them on the host
with its own generators and results are reproducible. Augmentations (train-time only) are out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr
from .engine import FusionEngine, _prefix

__all__ = ["build_samples", "generate_view_clip", "generate_view_clips"]


def _dev_tensor(x, dtype, dev):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return t.to(dev, dtype).contiguous()


@torch.no_grad()
def build_samples(samples: Sequence[Dict], view_ids: Sequence[Optional[Sequence[int]]], indices: Sequence, voxel_size: float,
                  use_color: bool = True, device="cuda") -> Dict:
    """`samples[i]`: dict with xyz (N,3), rgb (N,3), label (N,), per_obj (Q,C) and - unless the full cloud is
    used - vis_mask (V,N) (the arrays of the h5 groups `pointcloud/*` and `multiview/per_obj`,
    tools/preprocess_data.py:285-297). `view_ids[i]`: the views whose visibility is OR-ed
    (dataset_blender.py:338-346; None/empty = full cloud); `indices[i]`: the reference's
    `np.random.choice(np.arange(n_kept), MAX_POINTS, ...)` draw into the filtered cloud (:353-357).

    Returns the collated batch like `collate_fn` (coords (sum M', 4) int32 with the batch index first,
    input_features = [xyz, rgb] of the voxels, output_features = their target CLIP feature, labels int64,
    inverse_map per sample) plus the per-sample point-level arrays (xyz, rgb, feat, raw_label) under "points"."""
    eng = FusionEngine(device)
    lib, dev = eng.lib, eng.device
    n_s = len(samples)
    assert len(view_ids) == n_s and len(indices) == n_s
    n_pts = [int(np.shape(s["xyz"])[0]) for s in samples]
    n_obj = [int(s["per_obj"].shape[0]) for s in samples]
    dim = int(samples[0]["per_obj"].shape[1])
    lists = [list(map(int, v)) if v is not None else [] for v in view_ids]
    filtered = [len(v) > 0 for v in lists]
    n_rows = [int(len(ix)) for ix in indices]
    point_off, obj_off, out_off = _prefix(n_pts), _prefix(n_obj), _prefix(n_rows)
    vl_off = _prefix([len(v) for v in lists])
    n_views = [int(s["vis_mask"].shape[0]) if f else 0 for s, f in zip(samples, filtered)]
    for v, nv in zip(lists, n_views):
        if v and (min(v) < 0 or max(v) >= nv):
            raise IndexError(f"index {max(v)} is out of bounds for axis 0 with size {nv}")
    mask_off = _prefix([nv * n for nv, n in zip(n_views, n_pts)])
    up = lambda a: torch.from_numpy(a).to(dev)
    xyz = torch.cat([_dev_tensor(s["xyz"], torch.float64, dev).reshape(-1, 3) for s in samples])
    rgb = torch.cat([_dev_tensor(s["rgb"], torch.float64, dev).reshape(-1, 3) for s in samples])
    label = torch.cat([_dev_tensor(s["label"], torch.int64, dev).reshape(-1) for s in samples])
    per_obj = torch.cat([_dev_tensor(s["per_obj"], torch.float32, dev) for s in samples])
    masks = [_dev_tensor(np.asarray(s["vis_mask"]) != 0 if not isinstance(s["vis_mask"], torch.Tensor) else s["vis_mask"] != 0,
                         torch.uint8, dev).reshape(-1) for s, f in zip(samples, filtered) if f]
    vis = torch.cat(masks) if masks else torch.zeros(1, dtype=torch.uint8, device=dev)
    view_list = up(np.asarray([v for l in lists for v in l] or [0], dtype=np.int32))
    idx = torch.cat([_dev_tensor(np.asarray(ix, dtype=np.int64), torch.int64, dev).reshape(-1) for ix in indices]) \
        if sum(n_rows) else torch.zeros(1, dtype=torch.int64, device=dev)
    d_point_off, d_obj_off, d_out_off, d_vl_off, d_mask_off = up(point_off), up(obj_off), up(out_off), up(vl_off), up(mask_off)
    total, rows_total = int(point_off[-1]), int(out_off[-1])

    keep = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    check(lib.dc_sample_keep_flags(ptr(vis), ptr(d_mask_off), ptr(d_point_off), ptr(view_list), ptr(d_vl_off), n_s,
                                   max(n_pts, default=0), ptr(keep), current_stream()))
    new_index = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    kept_off = torch.empty(n_s + 1, dtype=torch.int64, device=dev)
    ws_bytes = lib.dc_compact_workspace(total)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.dc_compact_scan(ptr(keep), total, ptr(d_point_off), n_s, ptr(new_index), ptr(kept_off), ptr(ws), ws_bytes,
                              current_stream()))
    o_xyz = torch.empty((max(rows_total, 1), 3), dtype=torch.float32, device=dev)
    o_rgb = torch.empty((max(rows_total, 1), 3), dtype=torch.float32, device=dev)
    o_lab = torch.empty(max(rows_total, 1), dtype=torch.int32, device=dev)
    o_feat = torch.empty((max(rows_total, 1), dim), dtype=torch.float32, device=dev)
    rows = torch.empty(max(rows_total, 1), dtype=torch.int64, device=dev)
    kept_idx = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    mean = torch.empty(3 * n_s, dtype=torch.float64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    # label.astype(np.uint8) happens only on the filtered branch (dataset_blender.py:349); mixed batches are split
    as_u8 = filtered[0] if n_s else False
    if any(f != as_u8 for f in filtered):
        raise ValueError("build_samples: all samples of a call must either use view filtering or the full cloud")
    check(lib.dc_sample_gather(ptr(xyz), ptr(rgb), ptr(label), ptr(per_obj), ptr(d_obj_off), ptr(keep), ptr(new_index),
                               ptr(kept_off), ptr(d_point_off), ptr(idx), ptr(d_out_off), n_s, total, rows_total,
                               max(n_rows, default=0), dim, int(as_u8), ptr(o_xyz), ptr(o_rgb), ptr(o_lab), ptr(o_feat),
                               ptr(rows), ptr(kept_idx), ptr(mean), ptr(err), current_stream()))
    code = int(err.item())
    if code == 1:
        raise IndexError("index out of bounds: a point index exceeds the number of points visible in the chosen views")
    if code == 2:
        raise IndexError("index out of bounds: a label has no row in per_obj")
    o_xyz, o_rgb, o_lab, o_feat = o_xyz[:rows_total], o_rgb[:rows_total], o_lab[:rows_total], o_feat[:rows_total]

    # [feat, xyz, rgb] -> ME.utils.sparse_quantize(ignore_label=0, quantization_size=voxel_size)  (:400-414)
    cat = torch.cat([o_feat, o_xyz] + ([o_rgb] if use_color else []), dim=1)
    vox = eng.voxelize(o_xyz, d_out_off, float(voxel_size), o_lab, 0)
    voff = vox["voxel_off"].cpu().numpy()
    if (voff < 0).any():
        raise RuntimeError("sparse_quantize: a voxel coordinate fell outside [-2^20, 2^20)")
    vfeat = eng.voxel_gather(cat, d_out_off, vox, int(voff[-1]))
    coords, labels_v, inv = [], [], []
    for b in range(n_s):
        r0, m = int(out_off[b]), int(voff[b + 1] - voff[b])
        c = vox["coords"][r0:r0 + m]
        coords.append(torch.cat([torch.full((m, 1), b, dtype=torch.int32, device=dev), c], dim=1))
        labels_v.append(vox["voxel_labels"][r0:r0 + m].long())
        inv.append(vox["inverse_map"][r0:r0 + n_rows[b]])
    points = [{"xyz": o_xyz[out_off[b]:out_off[b + 1]], "rgb": o_rgb[out_off[b]:out_off[b + 1]],
               "feat": o_feat[out_off[b]:out_off[b + 1]], "raw_label": o_lab[out_off[b]:out_off[b + 1]]} for b in range(n_s)]
    return {"coords": torch.cat(coords) if coords else torch.zeros((0, 4), dtype=torch.int32, device=dev),
            "input_features": vfeat[:, dim:], "output_features": vfeat[:, :dim],
            "labels": torch.cat(labels_v) if labels_v else torch.zeros(0, dtype=torch.int64, device=dev),
            "inverse_map": inv, "voxel_off": voff, "points": points}


@torch.no_grad()
def generate_view_clips(pc, world_matrices, K, clip_features, h: int = 480, w: int = 640, device="cuda", return_device: bool = False):
    """Batched `MVDistilDataset.generate_view_clip` (data/dataset_blender.py:132-171): the per-point CLIP feature of
    each of V views of one scene. `pc` (N,3) float; `world_matrices` (V,4,4) camera->world as stored in
    `cameras.<scene>.json` (inverted in fp64 like utils/transforms.py:52-61); `K` (3,3) fp64 (`self.K`);
    `clip_features` (V, patch_h, patch_w, C) - the rearranged `CLIP.extract` output (:151). File reading and the
    CLIP tower stay with the caller. Every point is projected (truncation toward zero, z == 0 -> pixel (0,0)),
    the pixel is clipped into the image (:158-159, no visibility test) and the bicubically upsampled feature of
    that pixel is returned: (V, N, C) fp32 on the CPU like `.cpu()` (:170) unless `return_device`."""
    eng = FusionEngine(device)
    lib, dev = eng.lib, eng.device
    pts = _dev_tensor(np.asarray(pc) if not isinstance(pc, torch.Tensor) else pc, torch.float64, dev).reshape(-1, 3)
    poses = np.asarray(world_matrices, dtype=np.float64).reshape(-1, 4, 4)
    inv = np.ascontiguousarray(np.stack([np.linalg.inv(m) for m in poses]) if len(poses) else poses)
    feats = clip_features if isinstance(clip_features, torch.Tensor) else torch.from_numpy(np.asarray(clip_features))
    if feats.dim() == 3:
        feats = feats.unsqueeze(0)
    n_views, ph, pw, dim = (int(x) for x in feats.shape)
    if n_views != len(poses):
        raise ValueError(f"generate_view_clips: {len(poses)} camera poses for {n_views} feature maps")
    feats = feats.to(dev, torch.float32).contiguous()
    d_inv = torch.from_numpy(inv).to(dev)
    d_K = torch.from_numpy(np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))).to(dev)
    n = int(pts.shape[0])
    out = torch.empty((n_views, n, dim), dtype=torch.float32, device=dev)
    if n and n_views:
        check(lib.dc_view_clip_gather(ptr(pts), n, ptr(d_inv), ptr(d_K), ptr(feats), n_views, ph, pw, dim, int(h), int(w),
                                      ptr(out), current_stream()))
    if return_device:
        return out
    host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host


def generate_view_clip(pc, world_matrix, K, clip_feature, h: int = 480, w: int = 640, device="cuda"):
    """One view: (N, C) fp32 CPU tensor, the return value of the reference method (data/dataset_blender.py:132-171)."""
    feat = clip_feature if isinstance(clip_feature, torch.Tensor) else torch.from_numpy(np.asarray(clip_feature))
    return generate_view_clips(pc, np.asarray(world_matrix, dtype=np.float64)[None], K, feat[None], h, w, device)[0]
