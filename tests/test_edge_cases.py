"""Edge cases of the drop-in surface: empty clouds, clouds that no view sees, views that show only the table,
limits of the sorted visibility pipeline. Expectations come from the oracle restatement (which follows
utils/feature_fusion.py:272-343 line by line) on the same inputs."""
import numpy as np
import pytest
import torch

from tests import golden_io as gio

pytestmark = pytest.mark.gpu


def _mvff(sc, **kw):
    from dropclip_b200.feature_fusion import MultiviewFeatureFusion
    return MultiviewFeatureFusion(sc.intrinsic, image_size=(sc.intrinsic["height"], sc.intrinsic["width"]), device="cuda",
                                  use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False, **kw)


def _oracle(sc, points, colors, labels, segs, feats, return_obj=True):
    from oracle import fusion_ref
    K = fusion_ref.intrinsic_matrix(sc.intrinsic)
    return fusion_ref.fuse_object_level(points, colors, labels, sc.depths, segs, sc.camera_poses, feats, sc.query_embeddings, K,
                                        sc.intrinsic["height"], sc.intrinsic["width"], use_visibility=False, use_similarity=True,
                                        sim_method="max", return_obj=return_obj, device="cpu")


@pytest.mark.parametrize("return_obj", [True, False])
def test_cloud_that_no_view_sees(return_obj):
    sc = gio.scene_of(gio.load("fuse_s0.npz"))
    far = sc.points + 1000.0  # every point projects off-image or fails the depth test
    M = _mvff(sc)
    (f, w, vis), (p, c, l) = M.fuse(far, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                                    sc.query_embeddings, return_obj=return_obj, device="cuda")
    (of, ow, ovis), (op, oc, ol) = _oracle(sc, far, sc.colors, sc.labels, sc.seg_masks, sc.mv_features, return_obj)
    assert tuple(vis.shape) == tuple(ovis.shape) == (len(sc.depths), 0) and vis.dtype == torch.int64
    assert p.shape == op.shape == (0, 3) and c.shape == oc.shape and l.shape == ol.shape
    assert p.dtype == op.dtype and l.dtype == ol.dtype
    assert tuple(f.shape) == tuple(of.shape)
    if return_obj:  # the per-object features do not depend on the cloud at all
        g, o = f.cpu().numpy(), of.numpy()
        assert np.array_equal(np.isnan(g), np.isnan(o))
        ok = ~np.isnan(o)
        assert np.abs(g[ok] - o[ok]).max() <= 1e-3 * np.abs(o[ok]).max()
    assert np.allclose(w.cpu().numpy(), ow.numpy(), rtol=1e-3, atol=1e-6)


def test_empty_cloud_and_no_views():
    sc = gio.scene_of(gio.load("fuse_s0.npz"))
    M = _mvff(sc)
    m = M.get_visibility_mask(np.zeros((0, 3)), sc.depths, sc.camera_poses, device="cuda")
    assert tuple(m.shape) == (len(sc.depths), 0) and m.dtype == torch.int64 and m.device.type == "cpu"
    m = M.get_visibility_mask(sc.points, [], [], device="cuda")
    assert tuple(m.shape) == (0, sc.n_points) and m.dtype == torch.int64


def test_view_that_shows_only_the_table_contributes_nothing():
    sc = gio.scene_of(gio.load("fuse_s1.npz"))
    segs = [s.copy() for s in sc.seg_masks]
    segs[1][:] = 0  # np.unique(seg)[1:] == []: the rows of mv_features[1] are bound to no object
    M = _mvff(sc)
    (f, w, vis), _ = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses, sc.mv_features,
                            sc.query_embeddings, return_obj=True, device="cuda")
    (of, ow, ovis), _ = _oracle(sc, sc.points, sc.colors, sc.labels, segs, sc.mv_features)
    assert np.array_equal(vis.numpy(), ovis.numpy())
    assert (w.cpu().numpy()[:, 1] == 0).all() and np.allclose(w.cpu().numpy(), ow.numpy(), rtol=1e-3, atol=1e-6)
    g, o = f.cpu().numpy(), of.numpy()
    assert np.array_equal(np.isnan(g), np.isnan(o))
    ok = ~np.isnan(o)
    assert np.abs(g[ok] - o[ok]).max() <= 1e-3 * np.abs(o[ok]).max()


def test_sorted_visibility_limits_are_reported_not_truncated():
    from dropclip_b200 import _lib
    from dropclip_b200.engine import FusionEngine, SceneBatch
    sc = gio.scene_of(gio.load("fuse_s0.npz"))
    eng = FusionEngine("cuda")
    too_many = 820  # the constant-bank camera tables hold 819 views per scene
    scene = {"points": sc.points[:64], "depths": [sc.depths[0]] * too_many, "camera_poses": [sc.camera_poses[0]] * too_many,
             "intrinsic": sc.intrinsic}
    b = SceneBatch.from_host([scene], "cuda")
    with pytest.raises(_lib.DropClipError):
        eng.visibility_sorted(b, 0.05)
    mask, _, _ = eng.visibility(b, 0.05, torch.uint8)  # the literal kernel has no such limit
    assert mask.numel() == too_many * 64
    ok = dict(scene, depths=scene["depths"][:819], camera_poses=scene["camera_poses"][:819])
    b2 = SceneBatch.from_host([ok], "cuda")
    rec, rank, _ = eng.visibility_sorted(b2, 0.05)
    direct, _, _ = eng.visibility(b2, 0.05, torch.uint8)
    assert torch.equal(eng.unpack_visibility(b2, rec, rank, torch.uint8), direct)


def test_pixel_level_path_with_a_cloud_that_no_view_sees():
    from dropclip_b200.feature_fusion import MultiviewFeatureFusion
    z = gio.load("pixel_p0.npz")
    sc = gio.scene_of(z, pixel=True)
    M = MultiviewFeatureFusion(sc.intrinsic, image_size=(sc.intrinsic["height"], sc.intrinsic["width"]), device="cuda",
                               feature_size=int(sc.mv_features[0].shape[-1]), use_visibility=1, use_similarity=0,
                               use_obj_prior=0, norm_feat=True)
    far = sc.points + 1000.0
    (f, vis, sim), (p, c, l) = M.fuse(far, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                                      sc.query_embeddings, device="cuda")
    assert tuple(f.shape) == (0, sc.mv_features[0].shape[-1]) and tuple(vis.shape) == (len(sc.depths), 0) and sim is None
    assert p.shape == (0, 3) and c.shape[0] == 0 and l.shape == (0,)


def test_full_size_scene_properties():
    """One MV-TOD-sized scene (73 views, 480x640, 100 k points, 21 objects, 768-d): size-independent properties
    of the object-level pass - histogram totals and per-id counts, weights from the documented formula range,
    fused rows equal to the weighted mean of the bound feature rows recomputed in fp64, NaN rows exactly for
    objects bound in no view, compaction keeps exactly the points visible somewhere."""
    from dropclip_b200.engine import FusionEngine, batch_from_device
    from dropclip_b200.scenes import make_scene
    eng = FusionEngine("cuda")
    sc = make_scene(777, n_views=73, n_points=100_000, n_objects=21, device="cuda", as_torch=True)
    b = batch_from_device([sc], "cuda")
    res = eng.fuse_object_level(b, 0.05, False, True, "max", torch.uint8)
    counts, outside, row_object, object_row, status = eng.seg_tables(b)
    V, HW, Q, C = 73, 480 * 640, 21, 768
    assert int(status.abs().sum()) == 0 and int(outside.sum()) == 0
    assert torch.equal(counts.sum(1).long(), torch.full((V,), HW, device="cuda"))
    for v in (0, 36, 72):  # torch.bincount as an independent counter
        assert torch.equal(counts[v, :Q].long(), torch.bincount(b.segs[v].reshape(-1), minlength=Q)[:Q])
    w = res["weight_obj"][: Q * V].view(Q, V).double()
    rows = object_row[: Q * V].view(Q, V).long()
    assert (w[rows < 0] == 0).all() and (w[rows >= 0] >= 1e-6).all() and (w <= 1.0 + 1e-6).all()  # clip(pos - max neg, 1e-6), sims in [0,1]
    feats = b.feats.double()
    gathered = feats[rows.clamp(min=0)] * (rows >= 0).unsqueeze(-1)
    want = (gathered * w.unsqueeze(-1)).sum(1) / w.sum(1, keepdim=True)
    got = res["fused"].double()
    seen = w.sum(1) > 0
    assert torch.equal(torch.isnan(got).all(1), ~seen) and not bool(seen[0])  # the table (id 0) is never bound (quirk q10)
    rel = (got[seen] - want[seen]).abs().max() / want[seen].abs().max()
    assert float(rel) <= 1e-5
    new_index, kept_off, kept_host, out_off, cmask, _ = eng.compact_visibility(b, res["any_visible"], res["records"], res["rank"],
                                                                               torch.uint8)
    full = eng.unpack_visibility(b, res["records"], res["rank"], torch.uint8).view(V, -1)
    keep = full.sum(0) > 0
    assert torch.equal(keep, res["any_visible"].bool()) and int(kept_host[-1]) == int(keep.sum())
    assert torch.equal(cmask.view(V, -1), full[:, keep])
