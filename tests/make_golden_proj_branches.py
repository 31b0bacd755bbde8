"""Writes tests/golden/proj_branches.npz: outputs of the UNMODIFIED reference `project_2d_features_to_3d`
(utils/projections.py:108-147) for the branches the first golden file does not reach: center_crop (features aligned /
not aligned / image smaller than the crop -> torchvision pads), subsample_step, transform_coords None / regrad /
blender, with and without transform_to_world. Run in the build container only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "proj_branches.npz")
INTR = {"fx": 55.5, "fy": 54.5, "cx": 39.5, "cy": 29.5, "width": 80, "height": 60}

# name -> (depth shape, feature shape, kwargs); transform_coords given by name
CASES = {
    "crop_unaligned_feats": ((60, 80), (60, 80, 5), dict(center_crop=40, transform_coords="regrad", subsample_step=1)),
    "crop_aligned_feats": ((60, 80), (40, 40, 5), dict(center_crop=40, transform_coords="blender", subsample_step=1)),
    "crop_odd_margin": ((61, 79), (61, 79, 3), dict(center_crop=38, transform_coords=None, subsample_step=1)),
    "crop_pads_small_image": ((30, 80), (30, 80, 4), dict(center_crop=36, transform_coords="regrad", subsample_step=1)),
    "subsample_3": ((60, 80), (60, 80, 5), dict(transform_coords="regrad", subsample_step=3)),
    "subsample_none": ((60, 80), (60, 80, 5), dict(transform_coords="blender", subsample_step=None)),
    "crop_subsample_world": ((60, 80), (60, 80, 5), dict(center_crop=40, transform_coords="blender", subsample_step=7,
                                                         transform_to_world=True)),
    "subsample_world_nocvt": ((60, 80), (60, 80, 2), dict(transform_coords=None, subsample_step=5, transform_to_world=True)),
}


def case_inputs(name):
    dshape, fshape, kw = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % 1000 if False else sum(map(ord, name)))
    depth = rng.uniform(0.5, 4.0, size=dshape).astype(np.float32)
    feats = rng.standard_normal(fshape).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi)
    ext = np.eye(4, dtype=np.float32)
    ext[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=np.float32)
    ext[:3, 3] = rng.uniform(-1, 1, size=3)
    return depth, feats, ext, dict(kw)


def main():
    _, pj, _, _ = ref_shim.load()
    cvt = {"regrad": pj._cvt_regrad_coord, "blender": pj._cvt_blender_coord, None: None}
    g = {}
    for name in CASES:
        depth, feats, ext, kw = case_inputs(name)
        kw["transform_coords"] = cvt[kw["transform_coords"]]
        if kw.get("transform_to_world"):
            kw["camera_extrinsics"] = ext
        pc, f = pj.project_2d_features_to_3d(depth.copy(), feats.copy(), INTR, **kw)
        g[name + "_pc"], g[name + "_feat"] = np.asarray(pc), np.asarray(f)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
