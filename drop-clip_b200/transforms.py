"""Drop-in for the hot-path part of the reference's `utils/transforms.py`.

transform_pointcloud_to_world_frame / _to_camera_frame (utils/transforms.py:43-61) run on the GPU
with the reference's arithmetic (fp32 LAPACK inverse on the host, fp64 k-ascending FMA chain for
the 4x4 . 4xN product); CoordTransform2d (:99-146) and reconstruct_feature_map (:149-165) are
index arithmetic kept as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch


def transform_pointcloud_to_world_frame(pointcloud, camera_pose):
    from .projections import transform_points
    return transform_points(pointcloud, np.asarray(camera_pose))


def transform_pointcloud_to_camera_frame(pointcloud, camera_pose):
    from .projections import transform_points
    return transform_points(pointcloud, np.linalg.inv(camera_pose))


class CoordTransform2d:
    def __init__(self, img_dim, patch_size, resize_dim=None):
        self.height, self.width = img_dim
        self.crop_size = resize_dim or img_dim
        self.patch_size = patch_size
        self.patch_h = self.crop_size[0] / patch_size
        self.patch_w = self.crop_size[1] / patch_size

    @staticmethod
    def _transform(x, y, scale_h, scale_w):
        return (x * scale_w).long(), (y * scale_h).long()

    def img_to_patch(self, x, y):
        return self._transform(x, y, self.patch_h / self.height, self.patch_w / self.width)

    def patch_to_img(self, x, y):
        return self._transform(x, y, self.height / self.patch_h, self.width / self.patch_w)

    def crop_to_patch(self, x, y):
        return self._transform(x, y, self.patch_h / self.crop_size[0], self.patch_w / self.crop_size[1])

    def patch_to_crop(self, x, y):
        return self._transform(x, y, self.crop_size[0] / self.patch_h, self.crop_size[1] / self.patch_w)

    def img_to_crop(self, x, y):
        return self._transform(x, y, self.crop_size[0] / self.height, self.crop_size[1] / self.width)

    def crop_to_img(self, x, y):
        return self._transform(x, y, self.height / self.crop_size[0], self.width / self.crop_size[1])


def reconstruct_feature_map(feat, image_shape):
    """Nearest-patch upsampling feat[(y*ph/H).long(), (x*pw/W).long()] (utils/transforms.py:149-165)."""
    H, W, _ = image_shape
    ph, pw, _ = feat.shape
    y = torch.arange(H, device=feat.device).unsqueeze(1).expand(H, W).float()
    x = torch.arange(W, device=feat.device).unsqueeze(0).expand(H, W).float()
    return feat[(y * (ph / H)).long(), (x * (pw / W)).long()]
