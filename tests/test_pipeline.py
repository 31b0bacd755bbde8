"""FusionPipeline (pinned slots -> batched launch sequences -> async read-back) must hand out exactly what the
per-scene reference-shaped call returns (tools/preprocess_data.py:268: MVFF.fuse(..., return_obj=True))."""
import numpy as np
import pytest
import torch

from tests import golden_io as gio
from tests.test_gpu_parity import mvff, rel_close

pytestmark = pytest.mark.gpu


def _scenes():
    from dropclip_b200.scenes import small_scene
    out = []
    for i in range(7):  # ragged: different view / point / object counts, one scene nobody sees
        sc = small_scene(700 + i, n_views=3 + i % 4, n_points=1500 + 401 * i, n_objects=5 + i % 3, height=120, width=160)
        if i == 4:
            sc.points = sc.points + 1e4
        out.append(sc)
    return out


def _args(sc):
    return (sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features, sc.query_embeddings)


@pytest.mark.parametrize("batch", [1, 3])
def test_pipeline_equals_fuse(batch):
    from dropclip_b200.pipeline import FusionPipeline
    scs = _scenes()
    M = mvff(scs[0], use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False)
    want = [M.fuse(*_args(sc), return_obj=True, device="cuda") for sc in scs]
    pipe = FusionPipeline(scs[0].intrinsic, device="cuda", image_size=(120, 160), batch_scenes=batch, n_slots=4, max_views=6,
                          max_points=4000, max_queries=8)
    got = {}

    def consume():
        for r in pipe.results():
            got[r.tag] = r

    import threading
    t = threading.Thread(target=consume)
    t.start()
    for i, sc in enumerate(scs):
        slot = pipe.acquire()
        slot.fill(*_args(sc))
        pipe.submit(slot, tag=i)
    pipe.finish()
    t.join(60)
    pipe.close()
    assert sorted(got) == list(range(len(scs)))
    for i, ((f, w, vis), (p, c, l)) in enumerate(want):
        r = got[i]
        assert r.error is None
        assert np.array_equal(r.visibility_mask, vis.numpy().astype(np.uint8)) and r.visibility_mask.dtype == np.uint8
        assert np.array_equal(np.nan_to_num(r.mv_feats_obj, nan=5.0), np.nan_to_num(f.cpu().numpy(), nan=5.0))
        assert np.array_equal(r.weight_obj, w.cpu().numpy())
        rp, rc, rl = r.filtered()
        assert np.array_equal(rp, p) and np.array_equal(rc, c) and np.array_equal(rl, l)
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0 and pipe.launches > 0


def test_pipeline_reports_reference_errors_and_wide_ids():
    """An instance id >= Q is an IndexError in the reference (quirk q7): reported on that scene's result, the others are
    unaffected; a -1 background does not fit the uint8 slot and travels as int64."""
    from dropclip_b200.pipeline import FusionPipeline
    scs = _scenes()[:3]
    bad = scs[1]
    bad.seg_masks = [s.copy() for s in bad.seg_masks]
    bad.seg_masks[0][:4, :4] = 77
    neg = scs[2]
    neg.seg_masks = [s.copy() for s in neg.seg_masks]
    rng = np.random.default_rng(0)
    feats = []
    for v, s in enumerate(neg.seg_masks):
        s[:5, :] = -1
        ids = np.unique(s)[1:]
        feats.append(torch.from_numpy(rng.standard_normal((len(ids), 768)).astype(np.float32)).half())
    neg.mv_features = feats
    M = mvff(scs[0], use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False)
    (f2, w2, v2), _ = M.fuse(*_args(neg), return_obj=True, device="cuda")
    pipe = FusionPipeline(scs[0].intrinsic, device="cuda", image_size=(120, 160), batch_scenes=3, n_slots=3, max_views=6,
                          max_points=4000, max_queries=8)
    for i, sc in enumerate(scs):
        slot = pipe.acquire()
        slot.fill(*_args(sc))
        assert slot.wide_segs == (i == 2)
        pipe.submit(slot, tag=i)
    pipe.finish()
    res = {r.tag: r for r in pipe.results()}
    pipe.close()
    assert res[0].error is None and isinstance(res[1].error, IndexError) and res[2].error is None
    assert np.array_equal(res[2].weight_obj, w2.cpu().numpy())
    assert np.array_equal(res[2].visibility_mask, v2.numpy().astype(np.uint8))
