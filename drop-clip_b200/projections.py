"""Drop-in for the reference's `utils/projections.py`.

Same function names and argument meaning; geometry runs in libdropclip kernels (fp64, with the
reference's operation order). Functions whose arithmetic lives in Open3D / SciPy in the reference
(`rgbd_to_pointcloud_o3d`, the voxel down-sampling and KD-tree steps of `fuse_multiview_features*`)
are re-designed for the GPU (voxel hash + brute-force nearest neighbour) - see each docstring for
what is pinned against the reference and what is parity-unpinned.

Reference lines mirrored: utils/projections.py:16-261.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("dropclip_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda")


class CameraIntrinsics:
    def __init__(self, mat):
        self.fx = mat[0, 0]
        self.fy = mat[1, 1]
        self.cx = mat[0, 2]
        self.cy = mat[1, 2]

    def __iter__(self):
        return iter([self.fx, self.fy, self.cx, self.cy])

    @property
    def as_matrix(self):
        # the reference reads `self.xy` here (typo, utils/projections.py:31) and raises AttributeError
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1.0]])

    @property
    def as_dict(self):
        return {"fx": self.fx, "fy": self.fy, "cx": self.cx, "cy": self.cy}


def _k4(intr) -> torch.Tensor:
    return torch.tensor([intr["fx"], intr["fy"], intr["cx"], intr["cy"]], dtype=torch.float64, device=_dev())


def transform_points(pointcloud, matrix) -> np.ndarray:
    """(matrix . [p;1])[:3] in fp64, np.dot order (utils/transforms.py:43-61)."""
    lib = _lib.load()
    pts = torch.from_numpy(np.ascontiguousarray(pointcloud, dtype=np.float64).reshape(-1, 3)).to(_dev())
    out = torch.empty_like(pts)
    m = np.ascontiguousarray(np.asarray(matrix), dtype=np.float64).reshape(16)  # np.dot promotes to fp64 (points4 is fp64)
    check(lib.dc_transform_points(ptr(pts), pts.shape[0], m.ctypes.data_as(_lib.c_void_p), ptr(out), current_stream()))
    return out.cpu().numpy()


def pointcloud_to_pixel(pointcloud, camera_intrinsics):
    """x' = fx * x / z + cx, y' = fy * y / z + cy (un-truncated fp64), utils/projections.py:59-64."""
    lib = _lib.load()
    pts = torch.from_numpy(np.ascontiguousarray(pointcloud, dtype=np.float64).reshape(-1, 3)).to(_dev())
    out = torch.empty((pts.shape[0], 2), dtype=torch.float64, device=pts.device)
    check(lib.dc_points_to_pixels(ptr(pts), pts.shape[0], ptr(_k4(camera_intrinsics)), ptr(out), current_stream()))
    return out.cpu().numpy()


def backproject(depth_images, camera_intrinsics, flip_y=False, flip_z=False, poses=None, o3d_rounding=False) -> torch.Tensor:
    """Batched back-projection on the device: (V,H,W) fp32 -> (V,H,W,3) fp64 CUDA tensor."""
    lib = _lib.load()
    d = torch.as_tensor(np.ascontiguousarray(depth_images, dtype=np.float32)) if not isinstance(depth_images, torch.Tensor) \
        else depth_images.to(torch.float32)
    d = d.to(_dev()).contiguous()
    if d.dim() == 2:
        d = d.unsqueeze(0)
    V, H, W = d.shape
    out = torch.empty((V, H, W, 3), dtype=torch.float64, device=d.device)
    p = None
    if poses is not None:
        p = torch.from_numpy(np.ascontiguousarray(np.asarray(poses), dtype=np.float64).reshape(V, 16)).to(d.device)
    check(lib.dc_backproject(ptr(d), V, H, W, ptr(_k4(camera_intrinsics)), int(bool(flip_y)) | (2 if o3d_rounding else 0), int(flip_z), ptr(p), ptr(out),
                             current_stream()))
    return out


def depth_to_pointcloud(depth_image, camera_intrinsics):
    """(H,W) depth -> (H,W,3) camera-frame points, utils/projections.py:67-86. fp64 output like the
    reference's int-grid / python-float arithmetic."""
    depth = np.asarray(depth_image)
    if depth.dtype != np.float32:
        # the kernel takes fp32 depth (what the datasets store); wider inputs would lose bits
        if not np.array_equal(depth.astype(np.float32).astype(depth.dtype), depth):
            raise RuntimeError("depth_to_pointcloud: depth must be exactly representable in fp32")
    return backproject(depth, camera_intrinsics)[0].cpu().numpy()


def _cvt_regrad_coord(pts):
    pts[:, 2] = -pts[:, 2]
    pts[:, 1] = -pts[:, 1]
    return pts


def _cvt_blender_coord(pts):
    pts[:, 2] = -pts[:, 2]
    return pts


def apply_pca(features, norm=True, seed=42):
    """Visualisation helper (utils/projections.py:100-105); sklearn on the host, not on the hot path."""
    from sklearn.decomposition import PCA
    X = PCA(n_components=3, random_state=seed).fit_transform(features)
    if norm:
        X = (X - X.min()) / (X.max() - X.min())
    return X


def _center_crop(arr: np.ndarray, size: int) -> np.ndarray:
    """torchvision.transforms.CenterCrop([size, size]) on an (H, W[, C]) array, as utils/projections.py:119-125 applies it
    to the depth image (through PIL) and to the feature map: zero padding when the image is smaller than the crop
    (left/top get the floor half, right/bottom the ceil half), then the window at round((dim - size) / 2)."""
    h, w = arr.shape[:2]
    if size > w or size > h:
        pl, pt = ((size - w) // 2 if size > w else 0), ((size - h) // 2 if size > h else 0)
        pr, pb = ((size - w + 1) // 2 if size > w else 0), ((size - h + 1) // 2 if size > h else 0)
        arr = np.pad(arr, [(pt, pb), (pl, pr)] + [(0, 0)] * (arr.ndim - 2))
        h, w = arr.shape[:2]
        if (h, w) == (size, size):
            return np.ascontiguousarray(arr)
    top, left = int(round((h - size) / 2.0)), int(round((w - size) / 2.0))
    return np.ascontiguousarray(arr[top:top + size, left:left + size])


def project_2d_features_to_3d(depth_image, features, camera_intrinsics, center_crop=None, transform_to_world=False,
                              transform_coords=_cvt_regrad_coord, subsample_step=1, camera_extrinsics=None):
    """utils/projections.py:108-147: optional centre crop, back-projection, axis flip, strided
    sub-sampling, camera->world."""
    if center_crop:
        depth_image = _center_crop(np.asarray(depth_image), int(center_crop))
        if depth_image.shape[0:2] != features.shape[0:2]:  # crop features if not already aligned (:123-125)
            features = _center_crop(np.asarray(features), int(center_crop))
    pc = backproject(depth_image, camera_intrinsics)[0].reshape(-1, 3).cpu().numpy()
    features = features.reshape(-1, features.shape[-1])
    if transform_coords is not None:
        pc = transform_coords(pc)
    if subsample_step is not None:
        pc = pc[::subsample_step, ...]
        features = features[::subsample_step, ...]
    if transform_to_world:
        assert camera_extrinsics is not None
        pc = transform_points(pc, camera_extrinsics)
    return pc, features


def rgbd_to_pointcloud_o3d(rgb, depth, camera_intrinsics):
    """utils/projections.py:41-56 (depth_trunc=3 variant)."""
    from .geometry import rgbd_to_pointcloud_o3d as _impl
    return _impl(rgb, depth, camera_intrinsics, depth_scale=1.0, depth_trunc=3)


def pool_multiview_features(aggr_pc, aggr_features):
    """Unique points (np.unique(axis=0): lexicographic order) + per-point maximum of the features
    (utils/projections.py:245-261), as a GPU sort + segmented max. numpy in, numpy out."""
    lib = _lib.load()
    pts = torch.from_numpy(np.ascontiguousarray(aggr_pc, dtype=np.float64).reshape(-1, 3)).to(_dev())
    f_np = np.ascontiguousarray(aggr_features)
    in_dtype = f_np.dtype
    work = np.float64 if in_dtype == np.float64 else np.float32   # fp16/fp32 -> fp32 is exact for a max
    feats = torch.from_numpy(f_np.astype(work, copy=False).reshape(pts.shape[0], -1)).to(pts.device)
    n, dim = feats.shape
    out_p = torch.empty((max(n, 1), 3), dtype=torch.float64, device=pts.device)
    out_f = torch.empty((max(n, 1), dim), dtype=feats.dtype, device=pts.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=pts.device)
    ws = torch.empty(lib.dc_sort_workspace(n), dtype=torch.uint8, device=pts.device)
    check(lib.dc_unique_max_pool(ptr(pts), ptr(feats), _lib.DC_F64 if work == np.float64 else _lib.DC_F32, dim, n,
                                 ptr(out_p), ptr(out_f), ptr(cnt), ptr(ws), ws.numel(), current_stream()))
    u = int(cnt.item())
    return out_p[:u].cpu().numpy().astype(np.asarray(aggr_pc).dtype, copy=False), out_f[:u].cpu().numpy().astype(in_dtype, copy=False)


def _first_occurrence(nn: torch.Tensor, n_ref: int):
    """np.unique(nn, return_index=True) on the device: sorted unique values + first index of each."""
    first = torch.full((n_ref,), nn.numel(), dtype=torch.int64, device=nn.device)
    first.scatter_reduce_(0, nn, torch.arange(nn.numel(), device=nn.device), reduce="amin")
    ids = torch.nonzero(first < nn.numel()).squeeze(1)
    return ids, first[ids]


def fuse_multiview_features(pcs, multiview_features, camera_poses, camera_intrinsic, crop_size=336, patch_size=14,
                            voxel_size=0.0075, reshape_feat=False, norm_feat=True):
    """REGRAD-style pixel fusion (utils/projections.py:151-211): voxel-down the union cloud, map each
    view's points to their nearest union point, sample the nearest patch feature at the projected
    pixel of the first view point per union point, average over views. Returns
    (features (n,C) float64 tensor, pc_aggr numpy). Voxel down-sampling, nearest neighbour, rigid
    transform and projection run in libdropclip kernels; the remaining index bookkeeping uses
    torch indexing on the device (peripheral path)."""
    from .geometry import nearest_index, voxel_down
    dev = _dev()
    H, W = int(camera_intrinsic["height"]), int(camera_intrinsic["width"])
    pc_aggr = voxel_down(np.concatenate(pcs, axis=0), voxel_size)
    n = pc_aggr.shape[0]
    feats_all = multiview_features
    C = feats_all.shape[-1]
    ph = pw = crop_size // patch_size
    sums = torch.zeros((n, C), dtype=torch.float64, device=dev)
    counter = torch.zeros((n, 1), dtype=torch.float64, device=dev)
    for pc, feat, pose in zip(pcs, feats_all, camera_poses):
        nn = nearest_index(pc, pc_aggr)
        ids, first = _first_occurrence(nn, n)
        cam = transform_points(pc, np.linalg.inv(pose))
        px = torch.from_numpy(pointcloud_to_pixel(_cvt_regrad_coord(cam), camera_intrinsic)).to(dev)[first]
        if px.dim() < 2 or px.shape[0] == 0:
            continue
        pix = torch.where(torch.isfinite(px), px, torch.full_like(px, -9.3e18)).trunc().clamp_(-2 ** 62, 2 ** 62).long()
        ys = pix[:, 1].clamp(0, H - 1)
        xs = pix[:, 0].clamp(0, W - 1)
        if reshape_feat:
            feat = feat.reshape(ph, pw, C)
        if norm_feat:
            feat /= feat.norm(dim=-1, keepdim=True)  # in place, like the reference
        fdev = feat.to(dev)
        fh, fw = fdev.shape[0], fdev.shape[1]
        sel = fdev[(ys.float() * (fh / H)).long(), (xs.float() * (fw / W)).long()]
        sums[ids] = sums[ids] + sel.to(torch.float64)
        counter[ids] += 1
    counter[counter == 0] = 1e-5
    return (sums / counter).to(multiview_features.device), pc_aggr.cpu().numpy()


def fuse_multiview_features_obj_prior(pcs, pcs_label, multiview_features, obj_map, voxel_size=0.0075):
    """utils/projections.py:214-241: voxel-down the union cloud, transfer labels from the nearest raw
    point, unweighted mean over views of each object's feature, broadcast to the points (fp16)."""
    from .geometry import nearest_index, voxel_down
    dev = _dev()
    raw = np.concatenate(pcs, axis=0)
    raw_label = torch.from_numpy(np.concatenate(pcs_label, axis=0)).to(dev)
    pc_aggr = voxel_down(raw, voxel_size)
    label = raw_label[nearest_index(pc_aggr, raw)]
    feat_dev = multiview_features[0].device
    per_obj = torch.stack([torch.stack([f[i] for f in multiview_features], dim=0).mean(0) for i in range(len(obj_map))], dim=0)
    table = per_obj.to(dev, torch.float32)
    lut = torch.full((int(max(max(obj_map), int(label.max().item()) if label.numel() else 0)) + 2,), -1, dtype=torch.int64, device=dev)
    for i, o in enumerate(obj_map):
        lut[int(o)] = i
    rows = lut[label.clamp(min=0)]
    lib = _lib.load()
    q_off = torch.tensor([0, table.shape[0]], dtype=torch.int64, device=dev)
    p_off = torch.tensor([0, rows.numel()], dtype=torch.int64, device=dev)
    out = torch.empty((rows.numel(), table.shape[1]), dtype=torch.float32, device=dev)
    check(lib.dc_scatter_to_points(ptr(table.contiguous()), ptr(q_off), ptr(rows.contiguous()), ptr(p_off), 1, rows.numel(),
                                   int(table.shape[1]), 0, ptr(out), current_stream()))
    return out.to(torch.half).to(feat_dev), pc_aggr.cpu().numpy(), per_obj
