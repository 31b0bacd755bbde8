"""BASELINE configs[3]: pixel-level vs object-level fusion upper-bound ablation with the voxel-size sweep
(scripts/run_eval.py:120-286 driven by scripts/RUN_voxel_abls.bash: voxel_size in {0.002, 0.004, 0.006, 0.008} x world
scale), on one synthetic scene through the reference-shaped calls:

    aggregate_views_blender_new(scene, intrinsic, voxel_size)      utils/geometry.py:120-204 (run_eval.py:155)
    remove_table_mask                                              run_eval.py:156
    MVFF.fuse(... use_obj_prior=1, return_obj=True) -> feat[label] run_eval.py:240-251
    MVFF.fuse(... use_obj_prior=0)                                 run_eval.py:218
    CLIP.predict(mv_feats.half(), query, qneg=scene negatives, method=paired, threshold) per object   :270-277
    trainMetricPC(pred_list, gt_list)                              :286
and the training quantisation ME.utils.sparse_quantize(xyz, quantization_size=0.05) (data/dataset_blender.py:406-414)
of the fused cloud. The text tower is replaced by the scene's query embeddings (the positive prompt of object k is
its query row, the negatives are the other objects' rows - sim_negatives == 'scene')."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WORLD_SCALE = 10.0  # data/blender.py:183


def reference_scene_dict(sc):
    """The dict aggregate_views_blender_new walks: views[i] = {rgb, depth, annos=[(_, mask, colour)], camera}, col_to_ins."""
    rng = np.random.default_rng(0)
    n_obj = int(sc.query_embeddings.shape[0])
    col_to_ins = {i: i for i in range(n_obj)}
    views = {}
    for i, (d, s, p) in enumerate(zip(sc.depths, sc.seg_masks, sc.camera_poses)):
        ids = np.unique(s)
        annos = [(None, (s == k), int(k)) for k in ids]
        rgb = rng.integers(0, 256, size=s.shape + (3,), dtype=np.uint8)
        views[i] = {"rgb": rgb, "depth": d, "annos": annos, "camera": {"world_matrix": p}}
    return {"views": views, "col_to_ins": col_to_ins}


def _sync_ms(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def run(dev, n_views=8, n_objects=21, voxel_sizes=(0.002, 0.004, 0.006, 0.008), sim_thr=0.95, seed=1234):
    from dropclip_b200 import _lib
    from dropclip_b200.engine import FusionEngine
    from dropclip_b200.feature_fusion import MultiviewFeatureFusion
    from dropclip_b200.geometry import aggregate_views_blender_new, remove_table_mask
    from dropclip_b200.metrics import trainMetricPC
    from dropclip_b200.scenes import make_scene
    from dropclip_b200.voxelize import sparse_quantize
    eng = FusionEngine(dev)
    sc_obj = make_scene(seed, n_views=n_views, n_points=1000, n_objects=n_objects, device=str(dev))
    sc_pix = make_scene(seed, n_views=n_views, n_points=1000, n_objects=n_objects, device=str(dev), pixel_features=True,
                        feature_dtype=torch.float32)
    scene = reference_scene_dict(sc_obj)
    intr = sc_obj.intrinsic
    q = sc_obj.query_embeddings.to(dev)
    M_obj = MultiviewFeatureFusion(intr, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1,
                                   norm_feat=False, device=dev)
    M_pix = MultiviewFeatureFusion(intr, use_visibility=1, use_similarity=1, use_sim_kernel="max", use_obj_prior=0,
                                   norm_feat=True, device=dev)

    def ground(mv_feats, labels_dev):
        """run_eval.py:256-286 for one scene: per object (table skipped) paired prediction against the scene's other
        objects, then mean IoU / Pr@k."""
        preds, gts = [], []
        present = [int(k) for k in torch.unique(labels_dev).tolist() if k > 0]
        for k in present:
            text = torch.cat([q[k:k + 1], q[[j for j in range(1, n_objects) if j != k]]])
            score, pred = eng.predict(mv_feats.half(), text, _lib.DC_GROUND_PAIRED, 0.1, True, sim_thr)
            preds.append(pred.view(torch.bool))
            gts.append(labels_dev == k)
        iou, (p25, p50, p75) = trainMetricPC(preds, gts, pr_ious=[0.25, 0.5, 0.75], sigmoid=False)
        return float(iou), float(p25), float(p50), float(p75), len(present)

    rows = []
    for i, vs in enumerate((None,) + tuple(voxel_sizes)):  # a first untimed pass warms every kernel and allocator pool up
        vs_w = (voxel_sizes[0] if vs is None else vs) * WORLD_SCALE
        (points, colors, labels), t_aggr = _sync_ms(lambda: aggregate_views_blender_new(scene, intr, voxel_size=vs_w, depth_trunc=25.0))
        points, colors, labels = remove_table_mask(points, colors, labels)
        args = (points, colors, labels, sc_obj.depths, sc_obj.seg_masks, sc_obj.camera_poses)
        ((f_obj, w_obj, vis), (p2, c2, l2)), t_obj = _sync_ms(
            lambda: M_obj.fuse(*args, sc_obj.mv_features, sc_obj.query_embeddings, return_obj=True, device=dev))
        bad = torch.isnan(f_obj).any(1)
        f_obj = torch.where(bad[:, None], q, f_obj)                      # NaN rows <- query (run_eval.py:247-250)
        lab_dev = torch.from_numpy(np.ascontiguousarray(l2)).to(dev, torch.int64)
        mv_obj = f_obj[lab_dev]                                          # mv_feats_obj[labels] (:251)
        (g_obj), t_gobj = _sync_ms(lambda: ground(mv_obj, lab_dev))
        ((f_pix, vis_p, _), (p3, c3, l3)), t_pix = _sync_ms(
            lambda: M_pix.fuse(*args, [f.clone() for f in sc_pix.mv_features], sc_pix.query_embeddings, device=dev))
        lab3 = torch.from_numpy(np.ascontiguousarray(l3)).to(dev, torch.int64)
        f_pix = torch.nan_to_num(f_pix)
        (g_pix), t_gpix = _sync_ms(lambda: ground(f_pix, lab3))
        # training sample of data/dataset_blender.py:354-414: 10 000 random points, centre shift, quantise at 0.05 with the
        # (target feature, xyz, rgb) rows gathered per voxel and collided labels -> 0
        g = torch.Generator(device=dev).manual_seed(seed)
        n_fused = int(l2.shape[0])
        pick = torch.randperm(n_fused, generator=g, device=dev)[:10_000] if n_fused >= 10_000 else \
            torch.randint(0, max(n_fused, 1), (10_000,), generator=g, device=dev)
        xyz32 = torch.from_numpy(np.ascontiguousarray(p2, dtype=np.float32)).to(dev)[pick]
        rgb32 = torch.from_numpy(np.ascontiguousarray(c2, dtype=np.float32)).to(dev)[pick]
        xyz32 = xyz32 - xyz32.mean(0)
        cat = torch.cat([mv_obj[pick], xyz32, rgb32], dim=-1)
        (vox), t_vox = _sync_ms(lambda: sparse_quantize(xyz32, features=cat, labels=lab_dev[pick].int(), ignore_label=0,
                                                        return_index=True, return_inverse=True, quantization_size=0.05, device=dev))
        coords = vox[0]
        if vs is None:
            continue
        rows.append({"voxel_size": vs, "voxel_size_world": vs_w, "points": int(points.shape[0]), "points_visible": int(l2.shape[0]),
                     "aggregate_views_ms": t_aggr, "fuse_object_ms": t_obj, "fuse_pixel_ms": t_pix,
                     "ground_object_ms": t_gobj, "ground_pixel_ms": t_gpix, "objects_scored": g_obj[4],
                     "mIoU_object": g_obj[0], "Pr50_object": g_obj[2], "mIoU_pixel": g_pix[0], "Pr50_pixel": g_pix[2],
                     "train_voxels": int(coords.shape[0]), "train_quantize_ms": t_vox})
    return {"workload": "configs[3]: validate_upper_bound / run_eval shape, 1 scene, V=%d, 480x640, Q=%d, C=768; per voxel size: "
                        "aggregate_views_blender_new -> fuse (object level, pixel level) -> per-object paired predict (thr %.2f, "
                        "scene negatives) -> trainMetricPC; wall ms incl. host<->device copies of the reference-shaped calls"
                        % (n_views, n_objects, sim_thr), "sweep": rows}


if __name__ == "__main__":
    import json
    print(json.dumps(run(torch.device("cuda", 0)), indent=1))
