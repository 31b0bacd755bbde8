"""How far is the reference's own fp32 pixel-level result from the exact value of its formulas?

Runs the UNMODIFIED reference (utils/feature_fusion.py:138-270 through oracle/ref_shim.py; build container only) on the
configs[3] scene (V=8, 480x640, N=100k, Q=21, (24,32,768) fp32 patch maps, sim kernel max, norm_feat) with 1 and with
all BLAS/OpenMP threads, and oracle.fusion_ref with work=torch.float64 (the same formulas in double precision), and
prints the spread under the metric the parity tests use (|a - b| / max(|b|, 1e-3 max|b|)).

    python benchmarks/reference_fp32_noise.py > profiles/r02_reference_fp32_noise.md
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fusion_ref, ref_shim  # noqa: E402
from dropclip_b200.scenes import make_scene  # noqa: E402


def metric(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    ok = ~np.isnan(b)
    scale = np.maximum(np.abs(b), np.abs(b[ok]).max() * 1e-3)
    return np.where(ok, np.abs(a - b) / scale, 0.0)


def main():
    ff, _, _, _ = ref_shim.load()
    sc = make_scene(2234, n_views=8, n_points=100_000, n_objects=21, device="cpu", pixel_features=True, feature_dtype=torch.float32)
    segs = [torch.from_numpy(s) for s in sc.seg_masks]

    def reference(threads):
        torch.set_num_threads(threads)
        M = ff.MultiviewFeatureFusion(sc.intrinsic, use_visibility=1, use_similarity=1, use_sim_kernel="max", use_obj_prior=0,
                                      norm_feat=True, device="cpu")
        (f, v, w), _ = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses,
                              [x.clone() for x in sc.mv_features], sc.query_embeddings, device="cpu")
        return f.numpy(), w.numpy()

    n_thr = os.cpu_count() or 1
    f1, w1 = reference(1)
    fn, wn = reference(n_thr)
    K = fusion_ref.intrinsic_matrix(sc.intrinsic)
    (xf, _, xw), _ = fusion_ref.fuse_pixel_level(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses,
                                                 [x.clone() for x in sc.mv_features], sc.query_embeddings, K, 480, 640,
                                                 use_similarity=True, feature_size=768, sim_method="max", norm_feat=True,
                                                 work=torch.float64)
    xf, xw = xf.numpy(), xw.numpy()
    ef, ew = metric(fn, xf), metric(wn, xw)
    tf, tw = metric(f1, fn), metric(w1, wn)
    rows_bad = ef.max(1) > 1e-3
    wsum = xw.sum(0)
    print("# The reference's own fp32 noise on the pixel-level path (configs[3] scene, V=8, N'=%d kept points)\n" % xf.shape[0])
    print("Metric: |a - b| / max(|b|, 1e-3 max|b|), the one tests/test_gpu_parity.py::rel_close uses. `exact` = the same formulas")
    print("(utils/feature_fusion.py:138-270) evaluated in float64 by oracle/fusion_ref.py (work=torch.float64).\n")
    print("| comparison | similarity mask (8 x N') max | features (N' x 768) max | feature rows off by > 1e-3 |")
    print("|---|---|---|---|")
    print("| unmodified reference fp32 (%d threads) vs exact | %.3e | %.3e | %d of %d (%.2f %%) |" % (
        n_thr, ew.max(), ef.max(), rows_bad.sum(), rows_bad.size, 100.0 * rows_bad.mean()))
    print("| unmodified reference fp32, 1 thread vs %d threads | %.3e | %.3e | %d |" % (n_thr, tw.max(), tf.max(), (tf.max(1) > 1e-3).sum()))
    print()
    print("Rows off by more than 1e-3 have sum_v weight = %.2e (median; all rows: %.2e): they are the points whose views all give a"
          % (np.median(wsum[rows_bad]), np.median(wsum)))
    print("similarity weight pos - max(neg) next to the 1e-6 clip (two queries tie on that surface). The weight is then a difference of")
    print("two fp32 numbers ~1 carrying an absolute error ~1e-7, i.e. a relative error of several percent, and the fused feature is a")
    print("mean under such weights. No fp32 evaluation that is not the reference's exact instruction sequence (BLAS kernel, thread")
    print("partition, bicubic rounding) can reproduce those rows to 1e-3; the CUDA path therefore evaluates the similarity chain in")
    print("fp64 and the parity tests bound it by `exact_close`: within 1e-3 of the exact value everywhere, and within 1e-3 of the")
    print("reference up to the distance the reference itself keeps from the exact value on that element. No row is excluded.")


if __name__ == "__main__":
    main()
