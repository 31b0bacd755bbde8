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


def voxel_down_trace(points, colors, labels, voxel_size: float):
    """Open3D `voxel_down_sample_and_trace` followed by the label majority vote of
    aggregate_views_blender_new (utils/geometry.py:186-201), on the device: returns CUDA tensors
    (points (M,3) f64, colors (M,3) f64, labels (M,) i64, counts (M,) i64), voxels ordered by voxel index."""
    lib = _lib.load()
    pts = _f64(points)
    n = pts.shape[0]
    dev = pts.device
    cols = _f64(colors) if colors is not None else None
    labs = None
    if labels is not None:
        labs = (labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels))).to(dev, torch.int64)
        labs = labs.reshape(-1).contiguous()
    out_p = torch.empty((max(n, 1), 3), dtype=torch.float64, device=dev)
    out_c = torch.empty((max(n, 1), 3), dtype=torch.float64, device=dev) if cols is not None else None
    out_l = torch.empty(max(n, 1), dtype=torch.int64, device=dev) if labs is not None else None
    counts = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    m_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _workspace(n)
    check(lib.dc_voxel_down_trace(ptr(pts), ptr(cols), ptr(labs), n, float(voxel_size), ptr(out_p), ptr(out_c), ptr(out_l),
                                  None, ptr(counts), ptr(m_dev), ptr(ws), ws.numel(), current_stream()))
    m = int(m_dev.item())
    return out_p[:m], (out_c[:m] if out_c is not None else None), (out_l[:m] if out_l is not None else None), counts[:m]


def binary_masks_to_seg(masks, obj_ids=None):
    """utils/image.py:11-15 (host numpy; defines the instance-map dtype the hot path receives)."""
    if obj_ids is None:
        obj_ids = np.arange(masks.shape[0], dtype=np.uint8)
    return np.max(masks * obj_ids[:, None, None], axis=0)


def aggregate_views_blender_new(scene, camera_intrinsic, depth_trunc=25.0, voxel_size=None):
    """utils/geometry.py:120-204: back-project every view's valid pixels (depth < depth_trunc), flip to the
    Blender camera convention, move to the world frame, concatenate, and - with `voxel_size` - voxel-grid
    down-sample with mean position / colour and majority-vote instance label. Same arguments and return
    triple (numpy points, colors, labels) as the reference; all arithmetic on the device. Voxels come out
    ordered by voxel index (Open3D's order is its hash map's iteration order)."""
    dev = _dev()
    from .projections import backproject
    col_to_ins = scene["col_to_ins"]
    all_p, all_c, all_l = [], [], []
    for _, stuff in scene["views"].items():
        rgb = stuff["rgb"]
        depth = np.ascontiguousarray(stuff["depth"], dtype=np.float32)
        _, binary_masks, colors = zip(*stuff["annos"])
        seg = binary_masks_to_seg(np.stack(binary_masks), np.asarray([col_to_ins[x] for x in colors]))
        d = torch.from_numpy(depth).to(dev)
        valid_m = (d < float(depth_trunc)).reshape(-1)  # seg_ins_2d[valid_m]
        d0 = torch.where(d >= float(depth_trunc), torch.zeros_like(d), d)
        keep = (d0 > 0).reshape(-1)                     # Open3D keeps 0 < depth < trunc
        pts = backproject(d0, camera_intrinsic, o3d_rounding=True)[0].reshape(-1, 3)[keep]
        pts = (pts * torch.tensor([1.0, -1.0, -1.0], dtype=torch.float64, device=dev)).contiguous()  # pc.transform(T_cam)
        # camera -> world with the view's world_matrix as fp64 (`.astype(np.float64)`, utils/geometry.py:163-164)
        world = np.ascontiguousarray(np.asarray(stuff["camera"]["world_matrix"]), dtype=np.float64).reshape(16)
        out = torch.empty_like(pts)
        check(_lib.load().dc_transform_points(ptr(pts), pts.shape[0], world.ctypes.data_as(_lib.c_void_p), ptr(out),
                                              current_stream()))
        pts = out
        all_p.append(pts)
        all_c.append(torch.from_numpy(np.ascontiguousarray(rgb)).to(dev).reshape(-1, rgb.shape[-1])[keep].to(torch.float64))
        all_l.append(torch.from_numpy(np.ascontiguousarray(seg)).to(dev).reshape(-1)[valid_m].to(torch.int64))
    pts, cols, labs = torch.cat(all_p), torch.cat(all_c), torch.cat(all_l)
    if voxel_size is None:
        return pts.cpu().numpy(), cols.cpu().numpy() / 255.0, labs.cpu().numpy()
    # colours are divided by 255 before the down-sampling in Open3D (the cloud stores [0, 1] colours)
    cols = torch.from_numpy(cols.cpu().numpy() / 255.0).to(dev)
    p, c, l, _ = voxel_down_trace(pts, cols, labs, voxel_size)
    return p.cpu().numpy(), c.cpu().numpy(), l.cpu().numpy()


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
