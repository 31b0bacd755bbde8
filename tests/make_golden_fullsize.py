"""Golden vectors at the BASELINE.json config sizes, from the UNMODIFIED reference (imported from /root/reference
through oracle/ref_shim.py). Run here (build container) only:  python tests/make_golden_fullsize.py

  full_obj.npz    configs[1]: one scene, V=73 views of 480x640, N=100 000 points, Q=21, C=768 fp16 object rows,
                  production flags (tools/preprocess_data.py:177-185), MultiviewFeatureFusion.fuse(return_obj=True)
  full_pixel.npz  configs[3]: one scene, V=8, (24, 32, 768) fp32 patch maps, use_obj_prior=0, sim kernel max, norm_feat

The inputs (275 MB / 60 MB) are not stored: tests regenerate them with dropclip_b200.scenes.make_scene(seed, device="cpu")
(elementwise torch + numpy Generator streams) and compare SHA-256 digests of every input array with the ones stored
here before they trust the stored outputs. Outputs: the visibility masks as SHA-256 + per-view counts + the kept-point
bitmap; object features / weights in full (they are small); pixel-level features as 256 sampled rows in full plus the
per-row sum and L2 norm of every row, similarity weights for the sampled rows plus every column sum. The second file
also stores the same quantities evaluated in float64 by oracle.fusion_ref (work=torch.float64): the exact value of the
reference's formulas, which the tests need to tell the reference's own fp32 rounding noise from a real difference.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fusion_ref, ref_shim  # noqa: E402
from dropclip_b200.scenes import make_scene  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

OBJ_CASE = dict(seed=1234, n_views=73, n_points=100_000, n_objects=21)
PIX_CASE = dict(seed=2234, n_views=8, n_points=100_000, n_objects=21)
N_SAMPLED_ROWS = 256


def sha(a) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.view(np.uint8).reshape(-1).tobytes()).hexdigest()


def input_digests(sc) -> dict:
    feats = torch.cat([f.reshape(-1, f.shape[-1]) for f in sc.mv_features]).numpy()
    return {
        "in_points": sha(sc.points), "in_colors": sha(sc.colors), "in_labels": sha(sc.labels),
        "in_depths": sha(np.stack(sc.depths)), "in_segs": sha(np.stack(sc.seg_masks)),
        "in_poses": sha(np.stack(sc.camera_poses)), "in_feats": sha(feats), "in_query": sha(sc.query_embeddings.numpy()),
    }


def as_arrays(d: dict) -> dict:
    return {k: np.array(v) for k, v in d.items()}


def sampled_rows(n_rows: int) -> np.ndarray:
    return np.sort(np.random.default_rng(7).choice(n_rows, size=min(N_SAMPLED_ROWS, n_rows), replace=False))


def main():
    ff, _, _, _ = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)

    # ------------------------------------------------------------------ configs[1]: object level, V = 73
    t0 = time.time()
    sc = make_scene(device="cpu", **OBJ_CASE)
    M = ff.MultiviewFeatureFusion(sc.intrinsic, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1,
                                  norm_feat=False, device="cpu")
    (feat, w, vis), (p, c, l) = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses,
                                       sc.mv_features, sc.query_embeddings, return_obj=True, device="cpu")
    full = M.get_visibility_mask(sc.points, sc.depths, sc.camera_poses)
    keep = (full.sum(0) > 0).numpy()
    assert np.array_equal(full[:, keep].numpy(), vis.numpy())
    vis_u8 = vis.numpy().astype(np.uint8)
    out = dict(
        case=np.array([OBJ_CASE[k] for k in ("seed", "n_views", "n_points", "n_objects")]),
        kept=np.packbits(keep.astype(np.uint8)), n_kept=np.array([int(keep.sum())]),
        vis_sha=np.array(sha(vis_u8)), vis_full_sha=np.array(sha(full.numpy().astype(np.uint8))),
        vis_per_view=vis_u8.sum(1).astype(np.int64), vis_per_point_sha=np.array(sha(vis_u8.sum(0).astype(np.int32))),
        feat=feat.numpy(), weight=w.numpy(),
        kept_points_sha=np.array(sha(p)), kept_colors_sha=np.array(sha(c)), kept_labels_sha=np.array(sha(l)),
    )
    (pf, _, _), _ = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                           sc.query_embeddings, return_obj=False, device="cpu")
    out["point_feat_sha"] = np.array(sha(np.nan_to_num(pf.numpy(), nan=0.0)))  # rows are copies of `feat` rows (or zeros)
    out["point_feat_rows"] = pf.numpy()[sampled_rows(pf.shape[0])]
    np.savez_compressed(os.path.join(OUT, "full_obj.npz"), **as_arrays(input_digests(sc)), **out)
    print("full_obj: kept", int(keep.sum()), "visible pairs", int(vis_u8.sum()), "%.1f s" % (time.time() - t0))

    # ------------------------------------------------------------------ configs[3]: pixel level, V = 8
    t0 = time.time()
    sc = make_scene(device="cpu", pixel_features=True, feature_dtype=torch.float32, **PIX_CASE)
    M = ff.MultiviewFeatureFusion(sc.intrinsic, use_visibility=1, use_similarity=1, use_sim_kernel="max", use_obj_prior=0,
                                  norm_feat=True, device="cpu")
    segs = [torch.from_numpy(s) for s in sc.seg_masks]
    (feat, vis, simw), (p, c, l) = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses,
                                          [f.clone() for f in sc.mv_features], sc.query_embeddings, device="cpu")
    # the direct aggregate_features call (row a7): sums before the division, full-length masks
    sums, vis_all, simw_all = M.aggregate_features(sc.points, sc.depths, segs, sc.camera_poses,
                                                   [f.clone() for f in sc.mv_features], sc.query_embeddings, device="cpu")
    keep = (vis_all.sum(0) > 0).numpy()
    assert np.array_equal(vis_all[:, keep].numpy(), vis.numpy())
    K = fusion_ref.intrinsic_matrix(sc.intrinsic)
    (xf, xv, xw), _ = fusion_ref.fuse_pixel_level(
        sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses, [f.clone() for f in sc.mv_features],
        sc.query_embeddings, K, 480, 640, use_similarity=True, feature_size=768, sim_method="max", norm_feat=True,
        work=torch.float64)
    assert torch.equal(xv, vis)
    rows = sampled_rows(feat.shape[0])
    f32, f64 = feat.numpy(), xf.numpy()
    out = dict(
        case=np.array([PIX_CASE[k] for k in ("seed", "n_views", "n_points", "n_objects")]),
        kept=np.packbits(keep.astype(np.uint8)), n_kept=np.array([int(keep.sum())]),
        vis_sha=np.array(sha(vis.numpy().astype(np.uint8))), vis_per_view=vis.numpy().sum(1).astype(np.int64),
        rows=rows, feat_rows=f32[rows], feat_rows_exact=f64[rows].astype(np.float64),
        feat_row_sum=f32.astype(np.float64).sum(1).astype(np.float32), feat_row_norm=np.linalg.norm(f32.astype(np.float64), axis=1).astype(np.float32),
        feat_row_sum_exact=f64.sum(1), feat_row_norm_exact=np.linalg.norm(f64, axis=1),
        simw_rows=simw.numpy()[:, rows], simw_rows_exact=xw.numpy()[:, rows],
        simw_col_sum=simw.numpy().astype(np.float64).sum(0).astype(np.float32), simw_col_sum_exact=xw.numpy().sum(0),
        agg_sum_rows=sums.numpy()[np.flatnonzero(keep)[rows[:128]]], agg_simw_sha=np.array(sha(simw_all.numpy())),
        simw_sha=np.array(sha(simw.numpy())),
    )
    # how far the reference's own fp32 result is from the exact value of its formulas (same metric as the tests)
    nan = np.isnan(f64)
    scale = np.maximum(np.abs(f64), np.abs(f64[~nan]).max() * 1e-3)
    err = np.abs(f32 - f64) / scale
    out["ref_self_noise"] = np.array([float(np.nanmax(err)), float((np.nanmax(err, axis=1) > 1e-3).mean())])
    np.savez_compressed(os.path.join(OUT, "full_pixel.npz"), **as_arrays(input_digests(sc)), **out)
    print("full_pixel: kept", int(keep.sum()), "ref fp32 vs exact: max rel", out["ref_self_noise"][0],
          "rows off by > 1e-3:", out["ref_self_noise"][1], "%.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
