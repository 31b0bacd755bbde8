"""Synthetic MV-TOD-shaped tabletop scenes.

There is no dataset in this environment, so every test and benchmark input is an
analytic scene: a table plane z = 0 (instance id 0), `n_objects - 1` axis-aligned
boxes standing on it (ids 1..), `n_views` Blender-convention cameras on a hemisphere
looking at the origin (camera looks along -Z, +Y up, `world_matrix` is
camera->world, fp32 - the layout `data/blender.py:167-280` hands to the fusion path),
480x640 z-depth maps and int64 instance maps obtained by exact ray casting, and a
scene point cloud made of back-projected object pixels of all views, jittered and
shuffled (the reference's cloud comes out of an Open3D voxel hash map, i.e. in no
spatial order, `utils/geometry.py:120-204`).

Everything is elementwise torch so the same code runs on the CPU (tests, golden
vectors) and on the GPU (bench set-up, where 64 scenes x 73 views are needed fast).
No BLAS call is used: the generator must not depend on a matmul's summation order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

# data/blender.py:180-187 with base_scale = 10
MVTOD_INTRINSIC = {
    "height": 480,
    "width": 640,
    "fx": 444.44444444,
    "fy": 444.44444444,
    "cx": 319.5,
    "cy": 239.5,
}
BACKGROUND_DEPTH = 30.0  # > depth_trunc 25 used by tools/preprocess_data.py:222
TABLE_HALF_EXTENT = 6.0


def scaled_intrinsic(height: int, width: int) -> Dict[str, float]:
    """MV-TOD intrinsics rescaled to a smaller image (used by the small test scenes)."""
    s = height / 480.0
    return {
        "height": height,
        "width": width,
        "fx": MVTOD_INTRINSIC["fx"] * s,
        "fy": MVTOD_INTRINSIC["fy"] * s,
        "cx": (width - 1) / 2.0,
        "cy": (height - 1) / 2.0,
    }


@dataclass
class Scene:
    """One scene in exactly the containers `MultiviewFeatureFusion.fuse` receives."""

    points: np.ndarray  # (N,3) float64
    colors: np.ndarray  # (N,3) float64
    labels: np.ndarray  # (N,) int64
    depths: List[np.ndarray]  # V x (H,W) float32
    seg_masks: List[np.ndarray]  # V x (H,W) int64
    camera_poses: List[np.ndarray]  # V x (4,4) float32
    mv_features: List[torch.Tensor]  # V x (K_v, C) object rows, or V x (h, w, C) patch maps
    query_embeddings: torch.Tensor  # (Q, C) float32, rows L2-normalised
    intrinsic: Dict[str, float] = field(default_factory=dict)

    @property
    def n_views(self) -> int:
        return len(self.depths)

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])


def _look_at_pose(eye: torch.Tensor) -> torch.Tensor:
    """Camera->world matrices for cameras at `eye` (V,3) looking at the origin (fp64)."""
    fwd = -eye / eye.norm(dim=1, keepdim=True)  # viewing direction
    up = torch.tensor([0.0, 0.0, 1.0], dtype=eye.dtype, device=eye.device).expand_as(eye)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm(dim=1, keepdim=True)
    cam_up = torch.linalg.cross(right, fwd)
    pose = torch.zeros((eye.shape[0], 4, 4), dtype=eye.dtype, device=eye.device)
    pose[:, :3, 0] = right  # camera +X
    pose[:, :3, 1] = cam_up  # camera +Y
    pose[:, :3, 2] = -fwd  # camera +Z points backwards (Blender)
    pose[:, :3, 3] = eye
    pose[:, 3, 3] = 1.0
    return pose


def _sample_boxes(rng: np.random.Generator, n_boxes: int):
    """Non-overlapping axis-aligned boxes on the table: (n,3) min corner, (n,3) max corner."""
    lo, hi = [], []
    tries = 0
    while len(lo) < n_boxes:
        tries += 1
        side = rng.uniform(0.4, 1.2, size=2)
        height = rng.uniform(0.5, 2.0)
        c = rng.uniform(-4.0, 4.0, size=2)
        a = np.array([c[0] - side[0] / 2, c[1] - side[1] / 2, 0.0])
        b = np.array([c[0] + side[0] / 2, c[1] + side[1] / 2, height])
        ok = True
        if tries < 20000:  # after that accept overlaps instead of looping forever
            for a2, b2 in zip(lo, hi):
                if (a[0] < b2[0] + 0.1 and b[0] > a2[0] - 0.1 and a[1] < b2[1] + 0.1 and b[1] > a2[1] - 0.1):
                    ok = False
                    break
        if ok:
            lo.append(a)
            hi.append(b)
    return np.stack(lo), np.stack(hi)


def _render_views(pose64: torch.Tensor, box_lo: torch.Tensor, box_hi: torch.Tensor,
                  intr: Dict[str, float], chunk: int = 4):
    """Exact ray casting. Returns depth (V,H,W) fp32, seg (V,H,W) int64 and the fp64 rays."""
    dev = pose64.device
    H, W = int(intr["height"]), int(intr["width"])
    V = pose64.shape[0]
    us = torch.arange(W, dtype=torch.float64, device=dev)
    vs = torch.arange(H, dtype=torch.float64, device=dev)
    dx = ((us - intr["cx"]) / intr["fx"]).view(1, 1, W).expand(1, H, W)
    dy = ((vs - intr["cy"]) / intr["fy"]).view(1, H, 1).expand(1, H, W)
    depth = torch.empty((V, H, W), dtype=torch.float32, device=dev)
    seg = torch.empty((V, H, W), dtype=torch.int64, device=dev)
    for v0 in range(0, V, chunk):
        P = pose64[v0:v0 + chunk]
        n = P.shape[0]
        # direction in the flipped ("o3d") camera frame is (dx, dy, 1); Blender frame is (dx,-dy,-1)
        R = P[:, :3, :3]
        o = P[:, :3, 3]
        d = (R[:, :, 0].view(n, 1, 1, 3) * dx.unsqueeze(-1)
             - R[:, :, 1].view(n, 1, 1, 3) * dy.unsqueeze(-1)
             - R[:, :, 2].view(n, 1, 1, 3))  # (n,H,W,3) world direction, z' = 1 per unit t
        oz = o[:, 2].view(n, 1, 1)
        best_t = torch.full((n, H, W), float("inf"), dtype=torch.float64, device=dev)
        best_id = torch.zeros((n, H, W), dtype=torch.int64, device=dev)
        # table plane z = 0, finite extent
        dz = d[..., 2]
        t_pl = torch.where(dz < 0, -oz / dz, torch.full_like(dz, float("inf")))
        hx = o[:, 0].view(n, 1, 1) + t_pl * d[..., 0]
        hy = o[:, 1].view(n, 1, 1) + t_pl * d[..., 1]
        on_table = (hx.abs() <= TABLE_HALF_EXTENT) & (hy.abs() <= TABLE_HALF_EXTENT) & torch.isfinite(t_pl)
        best_t = torch.where(on_table, t_pl, best_t)
        inv_d = 1.0 / d
        for b in range(box_lo.shape[0]):
            t0 = (box_lo[b].view(1, 1, 1, 3) - o.view(n, 1, 1, 3)) * inv_d
            t1 = (box_hi[b].view(1, 1, 1, 3) - o.view(n, 1, 1, 3)) * inv_d
            tn = torch.minimum(t0, t1).amax(dim=-1)
            tf = torch.maximum(t0, t1).amin(dim=-1)
            hit = (tn <= tf) & (tn > 0) & (tn < best_t)
            best_t = torch.where(hit, tn, best_t)
            best_id = torch.where(hit, torch.full_like(best_id, b + 1), best_id)
        miss = ~torch.isfinite(best_t)
        best_t = torch.where(miss, torch.full_like(best_t, BACKGROUND_DEPTH), best_t)
        depth[v0:v0 + n] = best_t.to(torch.float32)
        seg[v0:v0 + n] = best_id
    return depth, seg


def make_scene(
    seed: int,
    n_views: int = 8,
    n_points: int = 100_000,
    n_objects: int = 21,
    intrinsic: Optional[Dict[str, float]] = None,
    feat_dim: int = 768,
    feature_dtype: torch.dtype = torch.float16,
    pixel_features: bool = False,
    patch_hw=(24, 32),
    point_jitter: float = 0.02,
    feature_noise: float = 1.0,
    device: str = "cpu",
    as_torch: bool = False,
):
    """Build one scene. `seed = 1234 + scene_idx` by convention.

    `as_torch=True` returns a dict of device tensors (stacked views) instead of a `Scene`
    of numpy arrays - used by the benchmark to keep set-up on the GPU.
    """
    intr = dict(intrinsic or MVTOD_INTRINSIC)
    H, W = int(intr["height"]), int(intr["width"])
    rng = np.random.default_rng(seed)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device(device)

    n_boxes = n_objects - 1
    lo_np, hi_np = _sample_boxes(rng, n_boxes)
    box_lo = torch.from_numpy(lo_np).to(dev)
    box_hi = torch.from_numpy(hi_np).to(dev)

    radius = rng.uniform(12.0, 16.0, size=n_views)
    azim = rng.uniform(0.0, 2 * math.pi, size=n_views)
    elev = rng.uniform(math.radians(20.0), math.radians(70.0), size=n_views)
    eye = np.stack([radius * np.cos(elev) * np.cos(azim),
                    radius * np.cos(elev) * np.sin(azim),
                    radius * np.sin(elev)], axis=1)
    pose64 = _look_at_pose(torch.from_numpy(eye).to(dev))
    pose32 = pose64.to(torch.float32)
    # render with the fp32-rounded pose: that is the matrix the fusion path is given
    depth, seg = _render_views(pose32.to(torch.float64), box_lo, box_hi, intr)

    # ---- scene point cloud: back-project object pixels of every view, jitter, shuffle
    per_view = -(-n_points // n_views)
    pts, labs = [], []
    P = pose32.to(torch.float64)
    for v in range(n_views):
        obj_px = torch.nonzero(seg[v].reshape(-1) > 0).squeeze(1)
        if obj_px.numel() == 0:
            obj_px = torch.arange(H * W, device=dev)
        pick = torch.randint(0, obj_px.numel(), (per_view,), generator=gen).to(dev)
        pix = obj_px[pick]
        pv = torch.div(pix, W, rounding_mode="floor")
        pu = pix - pv * W
        z = depth[v].reshape(-1)[pix].to(torch.float64)
        xc = (pu.to(torch.float64) - intr["cx"]) / intr["fx"] * z
        yc = (pv.to(torch.float64) - intr["cy"]) / intr["fy"] * z
        R, t = P[v, :3, :3], P[v, :3, 3]
        world = (R[:, 0].view(1, 3) * xc.view(-1, 1) - R[:, 1].view(1, 3) * yc.view(-1, 1)
                 - R[:, 2].view(1, 3) * z.view(-1, 1) + t.view(1, 3))
        pts.append(world)
        labs.append(seg[v].reshape(-1)[pix])
    pts = torch.cat(pts)[:n_points]
    labs = torch.cat(labs)[:n_points]
    jitter = torch.randn((n_points, 3), generator=gen, dtype=torch.float64).to(dev) * point_jitter
    pts = pts + jitter
    perm = torch.randperm(n_points, generator=gen).to(dev)
    pts, labs = pts[perm].contiguous(), labs[perm].contiguous()
    colors = torch.rand((n_points, 3), generator=gen, dtype=torch.float64).to(dev)

    # ---- CLIP-like embeddings
    q = torch.randn((n_objects, feat_dim), generator=gen, dtype=torch.float32)
    q = (q / q.norm(dim=-1, keepdim=True)).to(dev)
    mv_features = []
    if pixel_features:
        ph, pw = patch_hw
        for v in range(n_views):
            # nearest-patch object id -> query direction + noise
            ys = ((torch.arange(ph, device=dev).float() + 0.5) * (H / ph)).long().clamp_(0, H - 1)
            xs = ((torch.arange(pw, device=dev).float() + 0.5) * (W / pw)).long().clamp_(0, W - 1)
            ids = seg[v][ys][:, xs]
            noise = torch.randn((ph, pw, feat_dim), generator=gen, dtype=torch.float32).to(dev)
            f = q[ids] + feature_noise * noise / math.sqrt(feat_dim)
            mv_features.append(f.to(feature_dtype))
    else:
        for v in range(n_views):
            ids = torch.unique(seg[v])[1:]  # rows follow the sorted ids minus the smallest one
            noise = torch.randn((ids.numel(), feat_dim), generator=gen, dtype=torch.float32).to(dev)
            scale = 0.5 + 2.0 * torch.rand((ids.numel(), 1), generator=gen, dtype=torch.float32).to(dev)
            f = scale * (q[ids] + feature_noise * noise / math.sqrt(feat_dim))
            mv_features.append(f.to(feature_dtype))

    if as_torch:
        return {
            "points": pts, "colors": colors, "labels": labs, "depths": depth, "seg_masks": seg,
            "camera_poses": pose32, "mv_features": mv_features, "query_embeddings": q,
            "intrinsic": intr,
        }
    return Scene(
        points=pts.cpu().numpy(),
        colors=colors.cpu().numpy(),
        labels=labs.cpu().numpy(),
        depths=[d.cpu().numpy() for d in depth],
        seg_masks=[s.cpu().numpy() for s in seg],
        camera_poses=[p.cpu().numpy() for p in pose32],
        mv_features=[f.cpu() for f in mv_features],
        query_embeddings=q.cpu(),
        intrinsic=intr,
    )


def small_scene(seed: int, n_views: int = 4, n_points: int = 2000, n_objects: int = 6,
                height: int = 120, width: int = 160, **kw) -> Scene:
    """Quarter-resolution scene the CPU oracle fuses in well under a second."""
    return make_scene(seed, n_views=n_views, n_points=n_points, n_objects=n_objects,
                      intrinsic=scaled_intrinsic(height, width), **kw)
