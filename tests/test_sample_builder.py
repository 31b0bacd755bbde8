"""Training-sample assembly: CUDA batch builder vs the numpy restatement of
data/dataset_blender.py:330-362,400-414 with the same random draws."""
import numpy as np
import pytest
import torch


def _make(seed, n=5000, q=9, v=6, c=64):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-3, 3, size=(n, 3))
    rgb = rng.random((n, 3))
    label = rng.integers(0, q, size=n).astype(np.int64)
    per_obj = rng.standard_normal((q, c)).astype(np.float32)
    vis = (rng.random((v, n)) < 0.3).astype(np.int64)  # stored int64 by the reference (h5 vis_mask)
    return {"xyz": xyz, "rgb": rgb, "label": label, "per_obj": per_obj, "vis_mask": vis}


def test_oracle_sample_ref_shapes_and_center():
    from oracle import sample_ref
    s = _make(0)
    rng = np.random.default_rng(1)
    views = [1, 4]
    n_kept = int(s["vis_mask"][views].sum(0).astype(bool).sum())
    idx = rng.choice(np.arange(n_kept), 1000, replace=False)
    out = sample_ref.build_sample_ref(s["xyz"], s["rgb"], s["label"], s["per_obj"], s["vis_mask"], views, idx, 0.25)
    assert out["xyz"].shape == (1000, 3) and abs(out["xyz"].astype(np.float64).mean(0)).max() < 1e-6
    assert out["feat"].shape == (1000, 64) and out["coords"].dtype == np.int32
    assert np.array_equal(out["coords"][out["inverse_map"]], np.floor(out["xyz"] / np.float32(0.25)).astype(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("filtered", [True, False])
def test_build_samples_vs_restatement(filtered):
    from oracle import sample_ref
    from dropclip_b200.sample_builder import build_samples
    rng = np.random.default_rng(7)
    samples = [_make(10), _make(11, n=3000, q=5, v=3), _make(12, n=7001, q=21, v=8)]
    view_ids, indices = [], []
    for s in samples:
        v = list(rng.choice(s["vis_mask"].shape[0], size=2, replace=False)) if filtered else None
        n_kept = int(s["vis_mask"][v].sum(0).astype(bool).sum()) if filtered else s["xyz"].shape[0]
        m = 1500 if n_kept >= 1500 else n_kept + 40  # the second case draws with replacement, like the reference
        indices.append(rng.choice(np.arange(n_kept), m, replace=m > n_kept))
        view_ids.append(v)
    for use_color in (True, False):
        out = build_samples(samples, view_ids, indices, voxel_size=0.3, use_color=use_color)
        dim = 64
        for b, s in enumerate(samples):
            want = sample_ref.build_sample_ref(s["xyz"], s["rgb"], s["label"], s["per_obj"], s["vis_mask"], view_ids[b],
                                               indices[b], 0.3, use_color)
            pts = out["points"][b]
            assert np.array_equal(pts["xyz"].cpu().numpy(), want["xyz"])          # fp64 centring, sequential mean: bit-exact
            assert np.array_equal(pts["rgb"].cpu().numpy(), want["rgb"])
            assert np.array_equal(pts["feat"].cpu().numpy(), want["feat"])
            assert np.array_equal(pts["raw_label"].cpu().numpy(), want["raw_label"])
            v0, v1 = int(out["voxel_off"][b]), int(out["voxel_off"][b + 1])
            coords = out["coords"][v0:v1].cpu().numpy()
            assert (coords[:, 0] == b).all() and np.array_equal(coords[:, 1:], want["coords"])   # first-occurrence order
            assert np.array_equal(out["labels"][v0:v1].cpu().numpy(), want["vlabels"].astype(np.int64))
            assert np.array_equal(out["inverse_map"][b].cpu().numpy(), want["inverse_map"])
            assert np.array_equal(out["output_features"][v0:v1].cpu().numpy(), want["vfeat"][:, :dim])
            assert np.array_equal(out["input_features"][v0:v1].cpu().numpy(), want["vfeat"][:, dim:])
        assert out["input_features"].shape[1] == (6 if use_color else 3)


@pytest.mark.gpu
def test_build_samples_errors_like_numpy():
    from dropclip_b200.sample_builder import build_samples
    s = _make(3, n=500)
    with pytest.raises(IndexError):
        build_samples([s], [[0]], [np.array([10 ** 6])], 0.3)   # point index beyond the filtered cloud
    with pytest.raises(IndexError):
        build_samples([s], [[99]], [np.array([0])], 0.3)        # view id out of range
    bad = dict(s)
    bad["label"] = s["label"].copy()
    bad["label"][:] = 50                                       # no such per_obj row
    with pytest.raises(IndexError):
        build_samples([bad], [None], [np.arange(10)], 0.3)


# ---- generate_view_clip (data/dataset_blender.py:132-171), pinned by outputs of the unmodified reference method ----
VIEW_CLIP_CASES = ["scene", "wild", "small_odd_dim", "downsample"]


def _view_clip_case(name):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "view_clip.npz"))
    h, w = (int(x) for x in g[f"{name}_hw"])
    return g[f"{name}_pc"], g[f"{name}_world_matrix"], g[f"{name}_K"], torch.from_numpy(g[f"{name}_patch"]), h, w, g[f"{name}_out"]


@pytest.mark.parametrize("name", VIEW_CLIP_CASES)
def test_oracle_view_clip_vs_reference_golden(name):
    from oracle import sample_ref
    pc, wm, K, patch, h, w, want = _view_clip_case(name)
    got, pix = sample_ref.view_clip_ref(pc, wm, K, patch, h, w)
    assert got.dtype == torch.float32 and got.shape == want.shape
    np.testing.assert_array_equal(got.numpy(), want)  # same numpy/torch calls on the same machine: identical
    assert pix[:, 0].min() >= 0 and pix[:, 0].max() <= w - 1 and pix[:, 1].min() >= 0 and pix[:, 1].max() <= h - 1
    if name == "wild":
        assert (pix[:13] == 0).all()  # camera-plane points keep pixel (0,0); non-finite quotients clip to 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", VIEW_CLIP_CASES)
def test_generate_view_clip_vs_reference_golden(name):
    from dropclip_b200.sample_builder import generate_view_clip
    pc, wm, K, patch, h, w, want = _view_clip_case(name)
    got = generate_view_clip(pc, wm, K, patch, h, w)
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape and got.device.type == "cpu"
    # bicubic taps in fp32 in a different summation order than ATen: 1e-3 relative to the feature scale
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-3, atol=1e-3 * float(np.abs(want).max()))
    # every point must have landed on the same pixel: a wrong pixel shows up as an O(1) row difference
    row_err = np.abs(got.numpy() - want).max(1)
    assert row_err.max() < 5e-3 * float(np.abs(want).max())


@pytest.mark.gpu
def test_generate_view_clips_batch_equals_single_views_and_oracle():
    from oracle import sample_ref
    from dropclip_b200.sample_builder import generate_view_clip, generate_view_clips
    rng = np.random.default_rng(5)
    pc = rng.uniform(-3, 3, size=(1500, 3))
    K = np.array([[100.0, 0, 63.5], [0, 100.0, 47.5], [0, 0, 1]])
    poses = []
    for i in range(5):
        a = 2 * np.pi * i / 5
        m = np.eye(4)
        m[:3, 3] = [8 * np.cos(a), 8 * np.sin(a), 5.0]
        c, s = np.cos(a), np.sin(a)
        m[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) @ np.array([[0, 0, 1.0], [1.0, 0, 0], [0, 1.0, 0]])
        poses.append(m)
    patch = torch.from_numpy(rng.standard_normal((5, 12, 16, 64)).astype(np.float32))
    batch = generate_view_clips(pc, np.stack(poses), K, patch, 96, 128)
    assert tuple(batch.shape) == (5, 1500, 64)
    for v in range(5):
        single = generate_view_clip(pc, poses[v], K, patch[v], 96, 128)
        assert torch.equal(single, batch[v])
        want, _ = sample_ref.view_clip_ref(pc, poses[v], K, patch[v], 96, 128)
        np.testing.assert_allclose(batch[v].numpy(), want.numpy(), rtol=1e-3, atol=1e-3 * float(want.abs().max()))
    empty = generate_view_clips(np.zeros((0, 3)), np.stack(poses), K, patch, 96, 128)
    assert tuple(empty.shape) == (5, 0, 64)
