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
    m = np.ascontiguousarray(np.asarray(matrix), dtype=np.float32).reshape(16)
    check(lib.dc_transform_points(ptr(pts), pts.shape[0], m.ctypes.data_as(_lib.c_void_p), ptr(out), current_stream()))
    return out.cpu().numpy()


def pointcloud_to_pixel(pointcloud, camera_intrinsics):
    """x' = fx * x / z + cx, y' = fy * y / z + cy (un-truncated fp64), utils/projections.py:59-64."""
    lib = _lib.load()
    pts = torch.from_numpy(np.ascontiguousarray(pointcloud, dtype=np.float64).reshape(-1, 3)).to(_dev())
    out = torch.empty((pts.shape[0], 2), dtype=torch.float64, device=pts.device)
    check(lib.dc_points_to_pixels(ptr(pts), pts.shape[0], ptr(_k4(camera_intrinsics)), ptr(out), current_stream()))
    return out.cpu().numpy()


def backproject(depth_images, camera_intrinsics, flip_y=False, flip_z=False, poses=None) -> torch.Tensor:
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
        p = torch.from_numpy(np.ascontiguousarray(np.asarray(poses), dtype=np.float32).reshape(V, 16)).to(d.device)
    check(lib.dc_backproject(ptr(d), V, H, W, ptr(_k4(camera_intrinsics)), int(flip_y), int(flip_z), ptr(p), ptr(out),
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


def project_2d_features_to_3d(depth_image, features, camera_intrinsics, center_crop=None, transform_to_world=False,
                              transform_coords=_cvt_regrad_coord, subsample_step=1, camera_extrinsics=None):
    """utils/projections.py:108-147: optional centre crop, back-projection, axis flip, strided
    sub-sampling, camera->world."""
    if center_crop:
        h, w = depth_image.shape[:2]
        top, left = int(round((h - center_crop) / 2.0)), int(round((w - center_crop) / 2.0))  # torchvision CenterCrop
        depth_image = np.ascontiguousarray(depth_image[top:top + center_crop, left:left + center_crop])
        if depth_image.shape[0:2] != features.shape[0:2]:
            fh, fw = features.shape[:2]
            ft, fl = int(round((fh - center_crop) / 2.0)), int(round((fw - center_crop) / 2.0))
            features = features[ft:ft + center_crop, fl:fl + center_crop]
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
