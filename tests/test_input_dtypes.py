"""GPU tests for inputs the reference accepts in any dtype: point labels / colours / points of every fixed-width
dtype (utils/feature_fusion.py:277-281 just does `arr[mask]`, :133 compares `label == obj`), camera poses in fp64
(utils/transforms.py:54 inverts and multiplies in the pose's dtype) and instance maps with a negative background id
(np.unique(seg)[1:] drops it, :307)."""
import numpy as np
import pytest
import torch

from tests import golden_io as gio
from tests.test_gpu_parity import mvff, rel_close

pytestmark = pytest.mark.gpu


def _prod_flags(sc):
    return mvff(sc, use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False)


def _reference_point_feats(n_pts, labels, fused):
    """reconstruct_per_obj_feat (utils/feature_fusion.py:127-136) in numpy: rows of `fused` by `label == obj`, obj 0 skipped."""
    out = np.zeros((n_pts, fused.shape[1]), dtype=np.float32)
    for obj in range(1, fused.shape[0]):
        out[np.asarray(labels) == obj] = fused[obj]
    return out


@pytest.mark.parametrize("label_dtype", [np.int64, np.int32, np.int16, np.uint8, np.float32, np.float64])
def test_point_features_for_every_label_dtype(label_dtype):
    """fuse(..., return_obj=False) must scatter by the VALUE of the label whatever its dtype (the scatter kernel reads
    int64: feeding it the caller's int32 / uint8 / float rows would read out of bounds and fuse two labels into one id)."""
    z = gio.load("fuse_s1.npz")
    sc = gio.scene_of(z)
    M = _prod_flags(sc)
    labels = sc.labels.astype(label_dtype)
    if np.dtype(label_dtype).kind == "f":
        labels[::7] += 0.5  # non-integral float labels equal no object id -> zero rows in the reference
    (fo, _, _), _ = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                           sc.query_embeddings, return_obj=True, device="cuda")
    (pf, w, vis), (p, c, l) = M.fuse(sc.points, sc.colors, labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                                     sc.query_embeddings, return_obj=False, device="cuda")
    assert pf.device.type == "cpu" and pf.dtype == torch.float32
    assert l.dtype == np.dtype(label_dtype)
    keep = gio.unpack(z["kept"], sc.n_points).astype(bool)
    assert np.array_equal(l, labels[keep]) and np.array_equal(p, sc.points[keep])
    want = _reference_point_feats(l.shape[0], l, fo.cpu().numpy())
    got = pf.numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got, nan=7.0), np.nan_to_num(want, nan=7.0))  # pure row copies: exact
    # and against the reference's own run (row index of every point in the golden file) for integral labels
    if np.dtype(label_dtype).kind != "f":
        idx = z["point_feat_row"]
        ok = idx >= 0
        fo_np = fo.cpu().numpy()
        assert np.array_equal(np.nan_to_num(got[ok], nan=7.0), np.nan_to_num(fo_np[idx[ok]], nan=7.0))
        assert not got[~ok].any()


@pytest.mark.parametrize("kind", ["u8_colors_u8_labels", "f16_points_rows", "bool_labels_i16_colors", "odd_width_rows"])
def test_returned_rows_for_narrow_dtypes(kind):
    """Row widths that are not multiples of four bytes (uint8 RGB, uint8 / int16 labels, fp16 rows) are compacted on
    the device byte-granular; the reference's own h5 output stores labels as uint8 (tools/preprocess_data.py:294)."""
    z = gio.load("fuse_s0.npz")
    sc = gio.scene_of(z)
    M = _prod_flags(sc)
    rng = np.random.default_rng(0)
    n = sc.n_points
    points, colors, labels = sc.points, sc.colors, sc.labels
    if kind == "u8_colors_u8_labels":
        colors, labels = rng.integers(0, 256, size=(n, 3)).astype(np.uint8), sc.labels.astype(np.uint8)
    elif kind == "f16_points_rows":
        colors = rng.standard_normal((n, 3)).astype(np.float16)
    elif kind == "bool_labels_i16_colors":
        colors, labels = rng.integers(-9, 9, size=(n, 3)).astype(np.int16), (sc.labels > 0)
    else:
        colors = rng.integers(0, 256, size=(n, 5)).astype(np.uint8)  # 5-byte rows
    (f, w, vis), (p, c, l) = M.fuse(points, colors, labels, sc.depths, sc.seg_masks, sc.camera_poses, sc.mv_features,
                                    sc.query_embeddings, return_obj=True, device="cuda")
    keep = gio.unpack(z["kept"], n).astype(bool)
    assert c.dtype == colors.dtype and l.dtype == labels.dtype and p.dtype == points.dtype
    assert np.array_equal(p, points[keep]) and np.array_equal(c, colors[keep]) and np.array_equal(l, labels[keep])
    rel_close(f.cpu().numpy(), z["obj_sim_max_feat"], what="features unchanged by the row dtypes")


def test_fp64_poses_stay_fp64_end_to_end():
    """utils/transforms.py:54-58: np.linalg.inv and np.dot run in the POSE's dtype. An fp64 pose whose entries are
    not fp32-representable must be inverted and applied in fp64 (narrowing it flips mask bits near pixel borders)."""
    from dropclip_b200.engine import intrinsic_matrix
    from oracle import c_oracle
    sc, poses64, pts = gio.fp64_pose_case()
    sc.points = pts
    K = intrinsic_matrix(sc.intrinsic)
    want64 = c_oracle.visibility_mask(sc.points, sc.depths, poses64, K)
    want32 = c_oracle.visibility_mask(sc.points, sc.depths, [p.astype(np.float32) for p in poses64], K)
    assert (want64 != want32).sum() > 100, "the case must distinguish fp64 poses from their fp32 roundings"
    M = mvff(sc, use_similarity=False)
    got = M.get_visibility_mask(sc.points, sc.depths, poses64, device="cuda").numpy()
    assert np.array_equal(got, want64)
    got32 = M.get_visibility_mask(sc.points, sc.depths, [p.astype(np.float32) for p in poses64], device="cuda").numpy()
    assert np.array_equal(got32, want32)
    # literal fp64 kernel too
    from tests.test_gpu_parity import engine_visibility
    sc.camera_poses, sc.inv_poses = poses64, [np.linalg.inv(p) for p in poses64]
    direct, _, _ = engine_visibility(sc, sc.points, kernel="direct")
    assert np.array_equal(direct.astype(np.int64), want64)


def test_transform_helpers_keep_fp64_matrices():
    """transform_pointcloud_to_world_frame / _to_camera_frame with an fp64 matrix (utils/transforms.py:43-61)."""
    from dropclip_b200 import transforms
    rng = np.random.default_rng(3)
    pts = rng.uniform(-5, 5, size=(1000, 3))
    pose = np.eye(4)
    pose[:3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    pose[:3, 3] = rng.uniform(-3, 3, size=3)
    from oracle import c_oracle
    want = c_oracle.transform(pts, pose)  # k-ascending FMA chain of the BLAS dgemm, in fp64
    assert np.array_equal(want, np.dot(pose, np.vstack([pts.T, np.ones((1, pts.shape[0]))]))[:3, :].T), \
        "the C restatement must equal this host's np.dot (utils/transforms.py:45-47)"
    got = transforms.transform_pointcloud_to_world_frame(pts, pose)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    cam = transforms.transform_pointcloud_to_camera_frame(pts, pose)
    assert np.array_equal(cam, c_oracle.transform(pts, np.linalg.inv(pose)))
    # an fp32 pose: inverted in fp32 (utils/transforms.py:54), then promoted by np.dot
    p32 = pose.astype(np.float32)
    assert np.array_equal(transforms.transform_pointcloud_to_camera_frame(pts, p32), c_oracle.transform(pts, np.linalg.inv(p32)))


def test_negative_background_id_is_the_dropped_smallest_id():
    """np.unique(seg)[1:] (utils/feature_fusion.py:307) drops the smallest id whatever it is: with a -1 background the
    table (0) and every object keep their rows and nothing raises; the oracle runs the same inputs."""
    from dropclip_b200.engine import intrinsic_matrix
    from oracle import fusion_ref
    z = gio.load("fuse_s2.npz")
    sc = gio.scene_of(z)
    segs, feats = [], []
    rng = np.random.default_rng(5)
    for s, f in zip(sc.seg_masks, sc.mv_features):
        s = s.copy()
        s[:7, :] = -1  # a strip of background above the scene
        segs.append(s)
        ids = np.unique(s)[1:]
        feats.append(torch.from_numpy(rng.standard_normal((len(ids), 768)).astype(np.float32)).to(f.dtype))
    M = _prod_flags(sc)
    (f, w, vis), _ = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses, feats, sc.query_embeddings,
                            return_obj=True, device="cuda")
    H, W = sc.intrinsic["height"], sc.intrinsic["width"]
    (of, ow, ov), _ = fusion_ref.fuse_object_level(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses, feats,
                                                   sc.query_embeddings, intrinsic_matrix(sc.intrinsic), H, W, return_obj=True)
    assert np.array_equal(vis.numpy(), ov.numpy())
    rel_close(w.cpu().numpy(), ow.numpy(), what="weights with a -1 background")
    rel_close(f.cpu().numpy(), of.numpy(), what="features with a -1 background")
    assert (ow.numpy()[0] > 0).any(), "the table row must be bound now that -1 is the dropped id"
    # two distinct negative ids: the reference would index weight_obj[-1] (torch wraps); reported as IndexError here
    segs[0] = segs[0].copy()
    segs[0][:2, :] = -2
    feats[0] = torch.cat([feats[0][:1], feats[0]])
    with pytest.raises(IndexError):
        M.fuse(sc.points, sc.colors, sc.labels, sc.depths, segs, sc.camera_poses, feats, sc.query_embeddings,
               return_obj=True, device="cuda")
