"""ORACLE (test infrastructure, never the product path): CPU restatement of DROP-CLIP's
multi-view fusion in numpy / torch-CPU.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this module. It restates, function by function, what the reference does with
the same library calls (so that its wall time is representative of the reference's own CPU
torch path) but organised as free functions. Parity status: PINNED - every function below is
checked bit-for-bit (masks, assignments) or to 1e-6 (features) against the unmodified reference
imported from /root/reference by `tests/make_golden.py`; the resulting vectors are committed
under `tests/golden/` and re-checked by `tests/test_oracle_golden.py`.

Reference lines followed (all under /root/reference):
  camera_frame            utils/transforms.py:52-61, utils/feature_fusion.py:75-79
  project_to_pixels       utils/feature_fusion.py:90-104 (dup :201-211)
  view_visibility         utils/feature_fusion.py:105-123 (dup :212-229)
  visibility_mask         utils/feature_fusion.py:81-125
  relative_similarity     utils/feature_fusion.py:65-73
  fuse_object_level       utils/feature_fusion.py:272-343
  broadcast_object_feats  utils/feature_fusion.py:127-136
  aggregate_pixel_level   utils/feature_fusion.py:138-250
  fuse_pixel_level        utils/feature_fusion.py:252-270
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def intrinsic_matrix(intr: Dict[str, float]) -> np.ndarray:
    """utils/feature_fusion.py:35-40 - dtype follows the dict values (float -> fp64)."""
    return np.asarray([[intr["fx"], 0, intr["cx"]], [0, intr["fy"], intr["cy"]], [0, 0, 1]])


def camera_frame(points: np.ndarray, pose: np.ndarray) -> np.ndarray:
    """World -> flipped camera frame. inv() stays in the pose dtype (fp32 in production),
    the 4x4 . 4xN product is a BLAS dgemm when points are fp64, then y and z change sign."""
    inv_pose = np.linalg.inv(pose)
    homog = np.vstack([points.T, np.ones((1, points.shape[0]))])
    cam = np.dot(inv_pose, homog)[:3, :].T
    cam[:, 1] = -cam[:, 1]
    cam[:, 2] = -cam[:, 2]
    return cam


def project_to_pixels(points: np.ndarray, pose: np.ndarray, K: np.ndarray):
    """Returns (pixels (N,2) int64 [u,v], zdepth (N,) float). Division result is assigned
    into an int64 array, i.e. truncated toward zero; rows with z' == 0 keep pixel (0,0)."""
    n = points.shape[0]
    uvw = (K @ camera_frame(points, pose).T).T
    pix = np.zeros((n, 2), dtype=int)
    nz = uvw[:, 2] != 0
    with np.errstate(invalid="ignore", over="ignore"):
        pix[nz] = np.column_stack([[uvw[:, 0][nz] / uvw[:, 2][nz], uvw[:, 1][nz] / uvw[:, 2][nz]]]).T
    return pix, uvw[:, 2]


def view_visibility(pix: np.ndarray, zdepth: np.ndarray, depth: np.ndarray, height: int, width: int,
                    threshold: float, device="cpu"):
    """Bounds test + |sensor - z'| <= threshold (fp64 compare). Returns (bool (N,), pixels tensor)."""
    pt = torch.from_numpy(pix).to(device)
    sensor = torch.from_numpy(depth.copy()).to(device)
    zd = torch.from_numpy(zdepth).to(device)
    inside = (pt[:, 0] >= 0) * (pt[:, 1] >= 0) * (pt[:, 0] < width) * (pt[:, 1] < height)
    cols = pt.T
    ok = (torch.abs(sensor[cols[1][inside], cols[0][inside]] - zd[inside]) <= threshold).bool()
    inside[inside == True] = ok  # noqa: E712 - keeps the reference's masked assignment
    return inside, cols


def visibility_mask(points, depths, poses, K, height, width, threshold=0.05, device="cpu") -> torch.Tensor:
    """(V,N) int64 on the CPU regardless of `device` (reference quirk q5)."""
    out = torch.zeros((len(depths), points.shape[0]), dtype=int)
    for v, (depth, pose) in enumerate(zip(depths, poses)):
        pix, zd = project_to_pixels(points, pose, K)
        inside, _ = view_visibility(pix, zd, depth, height, width, threshold, device)
        out[v] = inside
    return out


def relative_similarity(pos, neg, method: str, eps: float = 1e-6, work=torch.float32):
    if method == "max":
        return torch.clip(pos - torch.max(neg, dim=-1)[0], eps).squeeze().to(work)
    if method == "mean":
        return torch.clip(pos - neg.mean(-1), eps).squeeze().to(work)
    raise ValueError("Please set method in [mean, max]")


def broadcast_object_feats(n_points: int, labels: np.ndarray, feat: torch.Tensor, obj_ids: Sequence[int]):
    out = torch.zeros((n_points, feat.shape[-1]), dtype=torch.float32)
    for i, obj in enumerate(obj_ids):
        if i == 0:
            continue
        out[np.argwhere(labels == obj), :] = feat[i]
    return out


def fuse_object_level(points, colors, labels, depths, seg_masks, poses, mv_features, query, K, height, width,
                      threshold=0.05, use_visibility=False, use_similarity=True, sim_method="max",
                      return_obj=False, device="cpu", work=torch.float32):
    """`work=torch.float32` is the reference's arithmetic. `work=torch.float64` evaluates the SAME formulas in
    double precision from the point where the reference goes to fp32 (`feat_v_norm.float() @ query.T`, :312: the
    normalised rows keep the feature dtype's own roundings, then everything is widened): the exact value the tests
    measure the reference's fp32 rounding noise against."""
    vis = visibility_mask(points, depths, poses, K, height, width, threshold, device)
    seen = (vis.sum(0) > 0).cpu().numpy()
    points, colors, labels = points[seen], colors[seen], labels[seen]
    vis = vis[:, seen]

    n_obj, n_views = query.shape[0], len(mv_features)
    stacked = torch.zeros((n_obj, n_views, 768), dtype=work, device=device)
    weight = torch.zeros((n_obj, n_views), dtype=work, device=device)
    for v in range(n_views):
        feat_v, seg = mv_features[v], seg_masks[v]
        ids = np.unique(seg)[1:].tolist()
        if use_similarity:
            if work == torch.float32:
                unit = feat_v / feat_v.norm(dim=-1, keepdim=True)
            else:
                # the correctly rounded norm: torch's fp32-accumulated norm of an fp16 row lands on the neighbouring
                # fp16 value for ~3.5 rows in 10 000 (whenever the exact norm is within 1e-7 of a rounding boundary)
                unit = feat_v / feat_v.double().norm(dim=-1, keepdim=True).to(feat_v.dtype)
            sim = unit.to(work) @ query.to(work).T
            sim = (sim - sim.min()) / (sim.max() - sim.min())
        for i, obj in enumerate(ids):
            weight[obj, v] = 1.0
            if use_visibility:
                weight[obj, v] = float((seg == obj).sum())
            if use_similarity:
                others = torch.as_tensor([o for o in range(n_obj) if o != obj]).long().to(device)
                weight[obj, v] = relative_similarity(sim[i][obj], sim[i][others], sim_method, work=work).item()
            stacked[obj, v] = feat_v[i]
    fused = torch.einsum("kvc,kv->kc", stacked, weight) / weight.sum(1).unsqueeze(-1)
    if not return_obj:
        feats = broadcast_object_feats(points.shape[0], labels, fused.float().cpu(), list(range(n_obj)))
    else:
        feats = fused
    return (feats, weight, vis), (points, colors, labels)


def aggregate_pixel_level(points, depths, seg_masks, poses, mv_features, query, K, height, width,
                          feature_size=768, threshold=0.05, use_similarity=True, sim_method="max",
                          norm_feat=True, device="cpu", work=torch.float32):
    """`work=torch.float32` is the reference's arithmetic. `work=torch.float64` evaluates the SAME formulas in
    double precision (features and queries widened first): the tests use it as the exact value when they
    measure how far the reference's own fp32 roundings are from it."""
    n, n_views = points.shape[0], len(depths)
    vis = torch.zeros((n_views, n), dtype=torch.long, device=device)
    simw = torch.zeros((n_views, n), dtype=work, device=device) if use_similarity else None
    acc = torch.zeros((n, feature_size), dtype=work, device=device)
    if work != torch.float32:
        mv_features = [f.to(work) for f in mv_features]
        query = query.to(work) if query is not None else None
    for v in range(n_views):
        fmap = F.interpolate(mv_features[v].permute(2, 0, 1).unsqueeze(0), size=(height, width),
                             mode="bicubic", align_corners=False).squeeze().permute(1, 2, 0)
        if norm_feat:
            fmap /= fmap.norm(dim=-1, keepdim=True)
        if use_similarity:
            raw = fmap.to(work) @ query.T
            metric = torch.zeros((height, width), dtype=work, device=device)
            for obj in range(len(query)):
                region = seg_masks[v] == obj
                sub = raw[region]
                others = torch.as_tensor([o for o in range(len(query)) if o != obj]).long().to(device)
                metric[region] = relative_similarity(sub[:, obj], sub[:, others], sim_method, work=work)
        pix, zd = project_to_pixels(points, poses[v], K)
        inside, cols = view_visibility(pix, zd, depths[v], height, width, threshold, device)
        vis[v] = inside
        sel = (inside == 1).nonzero(as_tuple=True)[0]
        xs, ys = cols[:, sel][0, :], cols[:, sel][1, :]
        if use_similarity:
            simw[v][inside == 1] = metric[ys, xs]
            contrib = fmap[ys, xs] * metric[ys, xs].unsqueeze(1)
        else:
            contrib = fmap[ys, xs]
        acc[inside == 1] += contrib
    return acc, vis, simw


def fuse_pixel_level(points, colors, labels, depths, seg_masks, poses, mv_features, query, K, height, width,
                     use_similarity=True, **kw):
    acc, vis, simw = aggregate_pixel_level(points, depths, seg_masks, poses, mv_features, query, K, height,
                                           width, use_similarity=use_similarity, **kw)
    seen = vis.sum(0) > 0
    keep = seen.cpu().numpy()
    points, colors, labels = points[keep], colors[keep], labels[keep]
    vis, acc = vis[:, seen], acc[seen]
    if use_similarity:
        simw = simw[:, seen]
    denom = vis.sum(0) if not use_similarity else simw.sum(0)
    acc /= denom.unsqueeze(1)
    return (acc, vis, simw), (points, colors, labels)
