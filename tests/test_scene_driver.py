"""shard.run_scene_driver = the scene loop of tools/preprocess_data.py:188-297 split over ranks (:704-730):
skip-if-exists, skip-if-missing, load into pinned slots, fuse, NaN rows <- query, one file per scene.

CPU part: the host logic (on-disk scene source, sharding, restart, writers, all_reduce of the statistics) with a
stand-in pipeline under gloo world_size 2. GPU part: the real FusionPipeline - every written file must hold exactly
what the per-scene reference-shaped call returns - and a two-rank launch."""
import os
import queue
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scenes(n=6):
    from dropclip_b200.scenes import small_scene
    return [small_scene(900 + i, n_views=3 + i % 3, n_points=1200 + 300 * i, n_objects=5 + i % 3, height=120, width=160)
            for i in range(n)]


def _save_all(root, scs, first_id=10, seg_dtype=np.uint8):
    from dropclip_b200.shard import SceneDirSource
    for i, sc in enumerate(scs):
        SceneDirSource.save(str(root), first_id + i, sc, seg_dtype=seg_dtype, objects_info="{%d: 'obj'}" % i)
    return SceneDirSource(str(root))


class _StandInPipeline:
    """Host-only stand-in with FusionPipeline's acquire/submit/release/finish/results surface: 'fuses' a scene into
    something every file field can be checked against (per_obj = per-view feature means, row 0 NaN)."""

    def __init__(self, n_slots=3):
        from dropclip_b200.pipeline import PinnedSceneSlot
        self._free = queue.Queue()
        for _ in range(n_slots):
            self._free.put(PinnedSceneSlot(120, 160, 6, 4000, 6 * 8, 8, pinned=False))
        self._results = queue.Queue()
        self.h2d_bytes = self.d2h_bytes = self.launches = 0

    def acquire(self):
        return self._free.get()

    def release(self, slot):
        self._free.put(slot)

    def submit(self, slot, tag=None):
        from dropclip_b200.pipeline import SceneResult
        V, N, Q = slot.n_views, slot.n_points, slot.n_queries
        per_obj = np.tile(slot.feats[:int(slot.feat_rows[:V].sum())].astype(np.float32).mean(0), (Q, 1))
        per_obj[0] = np.nan
        keep = np.arange(N) % 3 != 0
        vis = (slot.depths[:V, 0, :1] > 0).astype(np.uint8) * np.ones((V, int(keep.sum())), np.uint8)
        self.h2d_bytes += slot.input_bytes()
        self._results.put(SceneResult(tag, per_obj, np.ones((Q, V), np.float32), vis, keep,
                                      _src=(slot.points_src, slot.colors, slot.labels_src)))
        self._free.put(slot)

    def finish(self):
        self._results.put(None)

    def results(self):
        while True:
            r = self._results.get()
            if r is None:
                return
            yield r

    def close(self):
        pass


def test_scene_dir_source_roundtrip_reads_into_the_slot(tmp_path):
    from dropclip_b200.pipeline import PinnedSceneSlot
    scs = _scenes(2)
    src = _save_all(tmp_path, scs[:1], seg_dtype=np.uint8)
    _save_all(tmp_path, scs[1:], first_id=11, seg_dtype=np.int64)  # the reference's dtype: narrowed while loading
    assert src.ids() == [10, 11] and 10 in src and 12 not in src
    slot = PinnedSceneSlot(120, 160, 6, 4000, 48, 8, pinned=False)
    for sid, sc in zip((10, 11), scs):
        meta = src.load_into(sid, slot)
        V, N = sc.n_views, sc.n_points
        assert (slot.n_views, slot.n_points, slot.n_queries) == (V, N, sc.query_embeddings.shape[0])
        assert np.array_equal(slot.depths[:V], np.stack(sc.depths)) and not slot.wide_segs
        assert np.array_equal(slot.segs[:V], np.stack(sc.seg_masks).astype(np.uint8))
        assert np.array_equal(slot.points[:N], sc.points) and np.array_equal(slot.labels[:N], sc.labels)
        want_inv = np.stack([np.linalg.inv(p) for p in sc.camera_poses]).astype(np.float64).reshape(V, 16)
        assert np.array_equal(slot.inv_poses[:V], want_inv)
        assert np.array_equal(slot.feats[:sum(f.shape[0] for f in sc.mv_features)], torch.cat(sc.mv_features).numpy())
        assert list(slot.feat_rows[:V]) == [f.shape[0] for f in sc.mv_features]
        assert np.array_equal(meta["queries"], sc.query_embeddings.numpy()) and meta["objects_info"].startswith("{")
    # an id outside [0, 255] travels as int64 (np.unique()[1:] would drop a -1 background)
    neg = scs[0]
    neg.seg_masks = [m.copy() for m in neg.seg_masks]
    neg.seg_masks[0][:3] = -1
    _save_all(tmp_path, [neg], first_id=12, seg_dtype=np.int64)
    src.load_into(12, slot)
    assert slot.wide_segs and np.array_equal(slot.t_segs_wide.numpy()[:neg.n_views], np.stack(neg.seg_masks))


def test_rank_scene_ids_partitions():
    from dropclip_b200 import shard
    ids = [3, 4, 7, 9, 10, 11, 20]
    for split in ("strided", "reference"):
        parts = [shard.rank_scene_ids(ids, r, 3, split) for r in range(3)]
        assert sorted(sum(parts, [])) == ids
    assert shard.rank_scene_ids(ids, 1, 3, "strided") == [4, 10]
    assert shard.rank_scene_ids([], 0, 2) == []


def _check_files(out_dir, ids, scs_by_id, expect):
    from dropclip_b200 import shard
    for sid in ids:
        z = shard.read_scene(shard.output_path(str(out_dir), sid, "npz"))
        exp = expect(sid, scs_by_id[sid])
        for k, v in exp.items():
            assert z[k].dtype == v.dtype and z[k].shape == v.shape, (sid, k, z[k].dtype, z[k].shape, v.dtype, v.shape)
            assert np.array_equal(z[k], v, equal_nan=True), (sid, k)


def _standin_expect(sid, sc):
    Q, V, N = sc.query_embeddings.shape[0], sc.n_views, sc.n_points
    keep = np.arange(N) % 3 != 0
    per = np.tile(torch.cat(sc.mv_features).numpy().astype(np.float32).mean(0), (Q, 1))
    per[0] = sc.query_embeddings.numpy()[0]  # NaN row patched with the query (tools/preprocess_data.py:278-282)
    return {"multiview/per_obj": per, "multiview/obj_ids": np.arange(Q, dtype=np.uint8),
            "pointcloud/xyz": sc.points[keep].astype(np.float32), "pointcloud/rgb": sc.colors[keep].astype(np.float32),
            "pointcloud/label": sc.labels[keep].astype(np.uint8), "pointcloud/vis_mask": np.ones((V, int(keep.sum())), np.float32)}


def test_driver_host_logic_restart_and_missing(tmp_path):
    from dropclip_b200 import shard
    from dropclip_b200.scenes import scaled_intrinsic
    scs = _scenes(5)
    src = _save_all(tmp_path / "in", scs)
    out = tmp_path / "out"
    os.makedirs(out)
    open(out / "000011.h5py", "wb").close()  # a file the reference wrote: skipped (:192-195)
    st = shard.run_scene_driver(src, str(out), scaled_intrinsic(120, 160), scene_ids=[10, 11, 12, 13, 14, 15],
                                pipeline=_StandInPipeline(), loader_threads=2, writer_threads=2, fmt="npz")
    assert (st["assigned"], st["skipped_existing"], st["skipped_missing"], st["fused"], st["written"]) == (6, 1, 1, 4, 4)
    assert st["errors"] == [] and st["h2d_bytes"] > 0 and st["written_bytes"] > 0
    _check_files(out, [10, 12, 13, 14], {10 + i: s for i, s in enumerate(scs)}, _standin_expect)
    st2 = shard.run_scene_driver(src, str(out), scaled_intrinsic(120, 160), scene_ids=[10, 11, 12, 13, 14, 15],
                                 pipeline=_StandInPipeline(), fmt="npz")
    assert (st2["skipped_existing"], st2["fused"], st2["written"]) == (5, 0, 0)  # restart: nothing left to do


def _gloo_driver_worker(rank, world, port, in_dir, out_dir, q):
    import torch.distributed as dist
    from dropclip_b200 import shard
    from dropclip_b200.scenes import scaled_intrinsic
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        st = shard.run_scene_driver(shard.SceneDirSource(in_dir), out_dir, scaled_intrinsic(120, 160),
                                    pipeline=_StandInPipeline(), fmt="npz", loader_threads=2)
        q.put((rank, st["fused"], st["job"]))
    finally:
        dist.destroy_process_group()


def test_driver_two_ranks_on_gloo_write_disjoint_shares(tmp_path):
    import torch.multiprocessing as mp
    scs = _scenes(5)
    _save_all(tmp_path / "in", scs)
    out = tmp_path / "out"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_driver_worker, args=(r, 2, port, str(tmp_path / "in"), str(out), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [g[1] for g in got] == [3, 2]  # strided: rank 0 <- ids 10, 12, 14; rank 1 <- 11, 13
    assert got[0][2]["fused"] == 5 and got[0][2]["written"] == 5 and got[0][2] == got[1][2]
    _check_files(out, range(10, 15), {10 + i: s for i, s in enumerate(scs)}, _standin_expect)


# ------------------------------------------------------------------------------------------------------------ GPU
def _fuse_expect(M):
    def expect(sid, sc):
        (f, w, vis), (p, c, l) = M.fuse(sc.points, sc.colors, sc.labels, sc.depths, sc.seg_masks, sc.camera_poses,
                                        sc.mv_features, sc.query_embeddings, return_obj=True, device="cuda")
        per = f.cpu().numpy().copy()
        bad = np.isnan(per).any(1)
        per[bad] = sc.query_embeddings.numpy()[bad]
        return {"multiview/per_obj": per, "multiview/obj_ids": np.arange(per.shape[0], dtype=np.uint8),
                "pointcloud/xyz": p.astype(np.float32), "pointcloud/rgb": c.astype(np.float32),
                "pointcloud/label": l.astype(np.uint8), "pointcloud/vis_mask": vis.numpy().astype(np.float32)}
    return expect


@pytest.mark.gpu
def test_driver_files_equal_per_scene_fuse(tmp_path):
    from dropclip_b200 import shard
    from tests.test_gpu_parity import mvff
    scs = _scenes(6)
    src = _save_all(tmp_path / "in", scs, seg_dtype=np.int64)
    st = shard.run_scene_driver(src, str(tmp_path / "out"), scs[0].intrinsic, device="cuda:0", batch_scenes=4, fmt="npz",
                                image_size=(120, 160), max_views=6, max_points=4000, max_queries=8, n_slots=6)
    assert (st["fused"], st["written"], st["errors"]) == (6, 6, []) and st["launches"] > 0 and st["d2h_bytes"] > 0
    M = mvff(scs[0], use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False)
    _check_files(tmp_path / "out", range(10, 16), {10 + i: s for i, s in enumerate(scs)}, _fuse_expect(M))


_RANK_SCRIPT = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from dropclip_b200 import shard
from dropclip_b200.scenes import scaled_intrinsic
n_gpu = torch.cuda.device_count()
local = int(os.environ["LOCAL_RANK"]) % n_gpu
torch.cuda.set_device(local)
dist.init_process_group("nccl" if n_gpu >= int(os.environ["WORLD_SIZE"]) else "gloo")
st = shard.run_scene_driver(shard.SceneDirSource({inp!r}), {out!r}, scaled_intrinsic(120, 160), device="cuda:%d" % local,
                            batch_scenes=2, fmt="npz", image_size=(120, 160), max_views=6, max_points=4000, max_queries=8,
                            n_slots=4, pin_cores=True)
if st["rank"] == 0:
    print("JOB " + json.dumps(st["job"]))
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_driver_two_ranks_launched_by_torchrun(tmp_path):
    """Two ranks (two GPUs when the box has them, else both on cuda:0 with gloo for the statistics): disjoint shares,
    every file equal to the per-scene call, and a second launch finds nothing left to do."""
    from dropclip_b200 import shard
    from tests.test_gpu_parity import mvff
    scs = _scenes(5)
    _save_all(tmp_path / "in", scs)
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT, inp=str(tmp_path / "in"), out=str(tmp_path / "out")))
    port = 32500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    for expect_fused in (5, 0):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        import json
        job = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("JOB ")][0][4:])
        assert job["assigned"] == 5 and job["fused"] == expect_fused and job["written"] == expect_fused
        assert job["skipped_existing"] == 5 - expect_fused
    M = mvff(scs[0], use_visibility=0, use_similarity=1, use_sim_kernel="max", use_obj_prior=1, norm_feat=False)
    _check_files(tmp_path / "out", range(10, 15), {10 + i: s for i, s in enumerate(scs)}, _fuse_expect(M))
