"""GPU versions of the Open3D / SciPy glue the fusion path leans on (`utils/geometry.py`).

The reference delegates these to un-vendored native libraries (open3d 0.15.2, scipy cKDTree), so
parity is anchored on their published semantics and checked against numpy/scipy restatements
(tests/test_gpu_parity.py) - "parity unpinned" for the Open3D ones.

Reference lines: pc_voxel_down utils/geometry.py:350-352, find_closest_indices :390-401,
rgbd_to_pointcloud_o3d :21-36 (and utils/projections.py:41-56), remove_table_mask :294-300.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("dropclip_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda")


def _f64(x) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return t.to(_dev(), torch.float64).reshape(-1, 3).contiguous()


def _workspace(n: int) -> torch.Tensor:
    return torch.empty(_lib.load().dc_sort_workspace(int(n)), dtype=torch.uint8, device=_dev())


def voxel_down(points, voxel_size: float, return_first_index: bool = False):
    """Open3D `voxel_down_sample`: one output point per occupied voxel = mean of its points.
    Returns a CUDA fp64 tensor (M,3) (voxels ordered by voxel index)."""
    lib = _lib.load()
    pts = _f64(points)
    n = pts.shape[0]
    out = torch.empty((max(n, 1), 3), dtype=torch.float64, device=pts.device)
    first = torch.empty(max(n, 1), dtype=torch.int64, device=pts.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=pts.device)
    ws = _workspace(n)
    check(lib.dc_voxel_down_mean(ptr(pts), n, float(voxel_size), ptr(out), ptr(first), ptr(cnt), ptr(ws), ws.numel(),
                                 current_stream()))
    m = int(cnt.item())
    return (out[:m], first[:m]) if return_first_index else out[:m]


def pc_voxel_down(pc, voxel_size=0.0075):
    return voxel_down(pc, voxel_size).cpu().numpy()


def nearest_index(query, ref, return_dist2: bool = False):
    lib = _lib.load()
    q, r = _f64(query), _f64(ref)
    out = torch.empty(q.shape[0], dtype=torch.int64, device=q.device)
    d2 = torch.empty(q.shape[0], dtype=torch.float64, device=q.device) if return_dist2 else None
    check(lib.dc_nearest_index(ptr(q), q.shape[0], ptr(r), r.shape[0], ptr(out), ptr(d2), current_stream()))
    return (out, d2) if return_dist2 else out


def find_closest_indices(full_pc, filtered_pc, eps=None):
    """For each point of `filtered_pc` the index of its nearest neighbour in `full_pc`."""
    if eps is None:
        return nearest_index(filtered_pc, full_pc).cpu().numpy()
    idx, d2 = nearest_index(filtered_pc, full_pc, True)
    idx, dist = idx.cpu().numpy(), np.sqrt(d2.cpu().numpy())
    return idx[np.argwhere(dist <= eps)]


@dataclass
class PointCloud:
    """Stand-in for the open3d.geometry.PointCloud the reference returns (points / colors arrays)."""
    points: np.ndarray
    colors: np.ndarray


def rgbd_to_pointcloud_o3d(rgb, depth, camera_intrinsics, depth_scale=1.0, depth_trunc=25.0):
    """Open3D create_from_rgbd_image semantics: pixels with 0 < depth/scale < trunc, row-major,
    x = (u - cx) * z / fx, y = (v - cy) * z / fy, colours / 255."""
    dev = _dev()
    d = torch.as_tensor(np.ascontiguousarray(depth, dtype=np.float32)).to(dev) / float(depth_scale)
    d = torch.where(d >= float(depth_trunc), torch.zeros_like(d), d)
    keep = (d > 0).reshape(-1)
    from .projections import backproject
    pts = backproject(d, camera_intrinsics, o3d_rounding=True)[0].reshape(-1, 3)[keep]
    col = torch.as_tensor(np.ascontiguousarray(rgb)).to(dev).reshape(-1, rgb.shape[-1])[keep]
    # true division on the host: torch's CUDA div-by-scalar multiplies by the reciprocal (1 ulp off)
    return PointCloud(points=pts.cpu().numpy(), colors=col.cpu().numpy().astype(np.float64) / 255.0)


def remove_table_mask(points, colors, labels):
    keep = labels > 0
    return points[keep], colors[keep], labels[keep]
