"""Scene-parallel sharding across the GPUs of one box.

Scenes are independent (no cross-scene state anywhere in utils/feature_fusion.py); the reference
already shards contiguous scene-id ranges over worker processes (tools/preprocess_data.py:704-730)
and evaluation over DistributedSampler ranks (tools/train_distil.py:172). Here: one process per
GPU, rank r owns scenes r::world (or a cost-balanced assignment), the fusion kernels never
communicate, and torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only
to gather per-scene object features and to reduce metric sums - the same two collectives the
reference's eval loop performs (engine/distil.py:307-309, :475-493).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def strided_shard(n_scenes: int, rank: int, world: int) -> List[int]:
    """rank r <- scenes r, r + world, r + 2 world, ..."""
    return list(range(rank, n_scenes, world))


def reference_chunks(start: int, end: int, world: int) -> List[tuple]:
    """The reference's own split (tools/preprocess_data.py:711-716): chunk = ceil((end - start) / world) and
    worker n gets the INCLUSIVE id range [start + chunk * n, min(start + chunk * (n + 1), end)] (its scene loop
    runs range(start, 1 + end), :188). Adjacent ranges therefore share their boundary id; in the reference the
    second worker to reach it skips it because the file exists (:192-195)."""
    import math
    chunk = math.ceil((end - start) / world) if world > 0 else 0
    return [(start + chunk * n, min(start + chunk * (n + 1), end)) for n in range(world)]


def contiguous_shard(start: int, end: int, rank: int, world: int) -> List[int]:
    """Scene ids of `rank` under the reference's ceil-sized contiguous chunks (reference_chunks). The boundary id
    two neighbouring chunks share is owned by the lower rank only, so the ranks' lists are disjoint and their
    union is exactly [start, end] - the set of files the reference's workers produce."""
    lo, hi = reference_chunks(start, end, world)[rank]
    first = lo if rank == 0 else lo + 1
    return list(range(first, hi + 1))


def balanced_shard(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time assignment by cost (e.g. N * V per scene): deterministic, every rank
    computes the same table."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def gather_object_features(local_ids: Sequence[int], local_feats: Sequence[torch.Tensor], q_max: int,
                           device=None) -> Dict[int, torch.Tensor]:
    """All ranks end up with {scene_id: (Q_s, C) fp32} for every scene of the job. Payload per
    scene is padded to (q_max, C) (64.5 KB at Q=21, C=768); ranks may hold different scene counts."""
    rank, world = world_info()
    dim = int(local_feats[0].shape[1]) if len(local_feats) else 0
    device = device or (local_feats[0].device if len(local_feats) else torch.device("cpu"))
    n_local = torch.tensor([len(local_ids), dim], dtype=torch.int64, device=device)
    if world == 1:
        return {int(i): f for i, f in zip(local_ids, local_feats)}
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    n_max = max(int(c[0]) for c in counts)
    dim = max(int(c[1]) for c in counts)
    meta = torch.full((n_max, 2), -1, dtype=torch.int64, device=device)  # (scene id, Q_s)
    pay = torch.zeros((n_max, q_max, dim), dtype=torch.float32, device=device)
    for k, (i, f) in enumerate(zip(local_ids, local_feats)):
        meta[k, 0], meta[k, 1] = int(i), int(f.shape[0])
        pay[k, : f.shape[0]] = f.to(device=device, dtype=torch.float32)
    metas = [torch.empty_like(meta) for _ in range(world)]
    pays = [torch.empty_like(pay) for _ in range(world)]
    dist.all_gather(metas, meta)
    dist.all_gather(pays, pay)
    out: Dict[int, torch.Tensor] = {}
    for m, p in zip(metas, pays):
        for k in range(m.shape[0]):
            sid, q = int(m[k, 0]), int(m[k, 1])
            if sid >= 0:
                out[sid] = p[k, :q]
    return out


def reduce_metric_sums(values: torch.Tensor, average: bool = True) -> torch.Tensor:
    """all_reduce(SUM) then / world, as engine/distil.py:475-493 does for mIoU / Pr@k scalars."""
    rank, world = world_info()
    if world > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM)
        if average:
            values = values / world
    return values


def max_over_ranks(seconds: float, device=None) -> float:
    rank, world = world_info()
    if world == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _h5py():
    try:
        import h5py  # the reference's container; not installed in every environment
        return h5py
    except ImportError:
        return None


def output_path(out_dir: str, scene_id: int, fmt: Optional[str] = None) -> str:
    """`{id:06d}.h5py` (the reference's file name, tools/preprocess_data.py:191,285) when h5py is importable or
    fmt == "h5py"; `{id:06d}.npz` with the same group/dataset names otherwise."""
    fmt = fmt or ("h5py" if _h5py() is not None else "npz")
    return os.path.join(out_dir, "{:0>6}.{}".format(scene_id, fmt))


def scene_done(out_dir: str, scene_id: int) -> bool:
    """A scene counts as done if EITHER container exists - files the reference wrote (.h5py) are honoured."""
    return any(os.path.isfile(output_path(out_dir, scene_id, f)) for f in ("h5py", "npz"))


def pending_scenes(out_dir: str, scene_ids: Sequence[int]) -> List[int]:
    """Restart semantics of tools/preprocess_data.py:192-195: a scene whose output exists is skipped."""
    return [i for i in scene_ids if not scene_done(out_dir, i)]


def patch_nan_rows(per_obj: np.ndarray, query: np.ndarray) -> np.ndarray:
    """tools/preprocess_data.py:278-282: rows with any NaN (objects seen in no view, always the table row 0)
    are replaced by that object's query embedding. Returns a fresh fp32 array."""
    per_obj = np.array(per_obj, dtype=np.float32, copy=True)
    bad = np.isnan(per_obj).any(axis=1)
    per_obj[bad] = np.asarray(query, dtype=np.float32)[bad]
    return per_obj


def write_scene(out_dir: str, scene_id: int, per_obj: np.ndarray, query: np.ndarray, xyz, rgb, label, vis_mask,
                objects_info: Optional[str] = None, fmt: Optional[str] = None) -> str:
    """The reference's per-scene file (tools/preprocess_data.py:285-297): groups `multiview/{per_obj f32, obj_ids u8,
    objects_info str}` and `pointcloud/{xyz f32, rgb f32, label u8, vis_mask f32}`, NaN rows patched first
    (:278-282). Written as real HDF5 under the reference's `.h5py` name when h5py is importable (the reference's
    dataset loader reads it back unchanged); otherwise as `.npz` with "group/dataset" keys. The file appears
    atomically (rename), so a killed worker never leaves a half-written file that a restart would skip."""
    per_obj = patch_nan_rows(per_obj, query)
    os.makedirs(out_dir, exist_ok=True)
    h5 = _h5py() if fmt in (None, "h5py") else None
    if fmt == "h5py" and h5 is None:
        raise RuntimeError("write_scene(fmt='h5py') needs h5py")
    data = {
        "multiview/per_obj": per_obj,
        "multiview/obj_ids": np.arange(per_obj.shape[0], dtype=np.uint8),
        "pointcloud/xyz": np.asarray(xyz, dtype=np.float32),
        "pointcloud/rgb": np.asarray(rgb, dtype=np.float32),
        "pointcloud/label": np.asarray(label, dtype=np.uint8),
        "pointcloud/vis_mask": np.asarray(vis_mask, dtype=np.float32),
    }
    if h5 is not None:
        path = output_path(out_dir, scene_id, "h5py")
        tmp = path + ".tmp"
        with h5.File(tmp, "w") as hdf:
            mv, pc = hdf.create_group("multiview"), hdf.create_group("pointcloud")
            for key, arr in data.items():
                grp, name = key.split("/")
                (mv if grp == "multiview" else pc).create_dataset(name, data=arr, dtype=arr.dtype)
            mv.create_dataset("objects_info", data=str(objects_info or ""))
        os.replace(tmp, path)
        return path
    path = output_path(out_dir, scene_id, "npz")
    tmp = path + ".tmp.npz"
    np.savez(tmp, **data, **{"multiview/objects_info": np.asarray(objects_info or "")})
    os.replace(tmp, path)
    return path


def read_scene(path: str) -> Dict[str, np.ndarray]:
    """Reads either container back into {"group/dataset": array}."""
    if path.endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}
    h5 = _h5py()
    if h5 is None:
        raise RuntimeError("reading .h5py files needs h5py")
    out = {}
    with h5.File(path, "r") as hdf:
        for g in hdf:
            for k in hdf[g]:
                out[f"{g}/{k}"] = hdf[g][k][()]
    return out


# ---------------------------------------------------------------------------------------------- scene-parallel driver
# tools/preprocess_data.py:188-297 (scene loop) + :704-730 (split over workers), in its throughput form: the loop
# body of the reference is load -> fuse -> patch NaN rows -> write one file; here loader threads decode scenes
# straight into the pipeline's pinned slots, the GPU fuses batches of scenes, writer threads emit the files.

class SceneDirSource:
    """Scenes on disk, one directory `{id:06d}/` of plain `.npy` files per scene (what a dataset exporter writes once):
    depths (V,H,W) f32, segs (V,H,W) u8|i64, poses (V,4,4) f32|f64 camera->world, points (N,3) f64, colors (N,3),
    labels (N,), feats (sum K_v, C) f16|f32 with feat_rows (V,) i64, queries (Q,C) f32, optional objects_info.txt.
    `.npy` payloads are read with readinto() straight into the pinned slot arrays - no pageable staging copy."""

    FILES = ("depths", "segs", "poses", "points", "colors", "labels", "feats", "feat_rows", "queries")

    def __init__(self, root: str):
        self.root = root

    def scene_dir(self, scene_id: int) -> str:
        return os.path.join(self.root, "{:0>6}".format(scene_id))

    def ids(self) -> List[int]:
        out = []
        for name in sorted(os.listdir(self.root)):
            if name.isdigit() and os.path.isdir(os.path.join(self.root, name)):
                out.append(int(name))
        return out

    def __contains__(self, scene_id: int) -> bool:
        return os.path.isfile(os.path.join(self.scene_dir(scene_id), "depths.npy"))

    @staticmethod
    def save(root: str, scene_id: int, scene, seg_dtype=np.uint8, objects_info: Optional[str] = None) -> str:
        """Writes a scene given in fuse()'s containers (scenes.Scene or anything with the same attributes)."""
        d = os.path.join(root, "{:0>6}".format(scene_id))
        os.makedirs(d, exist_ok=True)
        feats = [f.detach().cpu() for f in scene.mv_features]
        arrs = {
            "depths": np.stack([np.asarray(x, dtype=np.float32) for x in scene.depths]),
            "segs": np.stack([np.asarray(x) for x in scene.seg_masks]).astype(seg_dtype),
            "poses": np.stack([np.asarray(p) for p in scene.camera_poses]),
            "points": np.asarray(scene.points, dtype=np.float64), "colors": np.asarray(scene.colors),
            "labels": np.asarray(scene.labels),
            "feats": torch.cat(feats).numpy() if feats else np.zeros((0, 768), np.float16),
            "feat_rows": np.asarray([int(f.shape[0]) for f in feats], dtype=np.int64),
            "queries": scene.query_embeddings.detach().cpu().to(torch.float32).numpy(),
        }
        for k, a in arrs.items():
            np.save(os.path.join(d, k + ".npy"), np.ascontiguousarray(a))
        if objects_info is not None:
            with open(os.path.join(d, "objects_info.txt"), "w") as f:
                f.write(objects_info)
        return d

    @staticmethod
    def _read_into(path: str, dst: Optional[np.ndarray], dtype=None):
        """Reads a .npy payload into `dst[:n]` (leading-axis prefix of a C-contiguous array of the file's dtype) with
        readinto(); returns the filled view. Without a destination (or on a dtype mismatch) returns np.load()."""
        with open(path, "rb") as f:
            major, minor = np.lib.format.read_magic(f)
            shape, fortran, dt = (np.lib.format.read_array_header_1_0 if major == 1 else np.lib.format.read_array_header_2_0)(f)
            if dst is None or fortran or dt != dst.dtype or tuple(shape[1:]) != tuple(dst.shape[1:]):
                f.seek(0)
                a = np.load(f, allow_pickle=False)
                return a if dtype is None else a.astype(dtype, copy=False)
            if shape[0] > dst.shape[0]:
                raise ValueError(f"{path}: {shape[0]} rows exceed the slot capacity {dst.shape[0]}")
            view = dst[:shape[0]]
            mv = memoryview(view).cast("B")
            got = 0
            while got < len(mv):
                n = f.readinto(mv[got:])
                if not n:
                    raise IOError(f"{path}: truncated payload")
                got += n
            return view

    def load_into(self, scene_id: int, slot) -> Dict[str, object]:
        """Fills a pipeline.PinnedSceneSlot from the scene's files; returns {"objects_info": str, "queries": (Q,C) f32}."""
        d = self.scene_dir(scene_id)
        p = lambda k: os.path.join(d, k + ".npy")
        depths = self._read_into(p("depths"), slot.depths)
        V = int(depths.shape[0])
        if not np.shares_memory(depths, slot.depths):  # dtype mismatch: np.load()ed copy
            slot.depths[:V] = depths
        segs = self._read_into(p("segs"), slot.segs)
        slot.wide_segs = False
        if not (isinstance(segs, np.ndarray) and segs.dtype == np.uint8 and np.shares_memory(segs, slot.segs)):
            slot.store_segs(segs)
        slot.set_poses(list(np.load(p("poses"))))
        points = np.load(p("points"))
        labels = np.load(p("labels"))
        colors = np.load(p("colors"))
        N = int(points.shape[0])
        feat_rows = np.load(p("feat_rows"))
        feats = self._read_into(p("feats"), slot.feats)
        if not np.shares_memory(feats, slot.feats):
            slot.t_feats[:feats.shape[0]].copy_(torch.from_numpy(feats))
        queries = self._read_into(p("queries"), slot.queries)
        if not np.shares_memory(queries, slot.queries):
            slot.queries[:queries.shape[0]] = queries
        Q = int(queries.shape[0])
        slot.set_scene(V, N, Q, feat_rows, points, colors, labels)
        info = ""
        ip = os.path.join(d, "objects_info.txt")
        if os.path.isfile(ip):
            with open(ip) as f:
                info = f.read()
        return {"objects_info": info, "queries": np.array(slot.queries[:Q], copy=True)}


def rank_scene_ids(scene_ids: Sequence[int], rank: int, world: int, split: str = "strided") -> List[int]:
    """Scene ids of one rank. "strided": r, r+world, ... over the sorted id list (balanced whatever the id gaps);
    "reference": the ceil-sized contiguous id ranges of tools/preprocess_data.py:711-716 (contiguous_shard)."""
    ids = sorted(int(i) for i in scene_ids)
    if not ids:
        return []
    if split == "strided":
        return [ids[k] for k in strided_shard(len(ids), rank, world)]
    if split == "reference":
        mine = set(contiguous_shard(ids[0], ids[-1], rank, world))
        return [i for i in ids if i in mine]
    raise ValueError("split must be 'strided' or 'reference'")


def _gpu_local_cores(device_index: int) -> List[int]:
    """CPUs NVML reports as local to the GPU (its NUMA node / PCIe root), [] when unknown."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            parts = visible.split(",")
            if device_index < len(parts) and parts[device_index].isdigit():
                idx = int(parts[device_index])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        return [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
    except Exception:
        return []


def pin_rank_cores(rank_local: Optional[int] = None, world_local: Optional[int] = None, device_index: Optional[int] = None) -> List[int]:
    """Gives each rank of a node its own slice of the host cores, taken from the cores local to ITS GPU when NVML knows
    them (pinned staging memory allocated afterwards then sits on the GPU's NUMA node, and the loader / staging / writer
    threads of different ranks never share a core). Without NVML affinity information: contiguous slices of the allowed
    cores by local rank. Returns the cores now allowed; no-op for one rank."""
    rank_local = int(os.environ.get("LOCAL_RANK", "0")) if rank_local is None else rank_local
    world_local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1) if world_local is None else world_local
    try:
        cores = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return []
    if world_local <= 1 or len(cores) < world_local:
        return cores
    mine: List[int] = []
    local = [c for c in _gpu_local_cores(rank_local if device_index is None else device_index) if c in set(cores)]
    if local and len(local) < len(cores):
        # the ranks whose GPUs share this locality set split it evenly, in local-rank order
        sharers = [r for r in range(world_local) if sorted(c for c in _gpu_local_cores(r) if c in set(cores)) == local]
        if rank_local in sharers and len(local) >= len(sharers):
            per = len(local) // len(sharers)
            k = sharers.index(rank_local)
            mine = local[k * per:(k + 1) * per]
    if not mine:
        per = len(cores) // world_local
        mine = cores[rank_local * per:(rank_local + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return cores
    return mine


def run_scene_driver(source, out_dir: str, camera_intrinsic: Dict[str, float], scene_ids: Optional[Sequence[int]] = None,
                     device=None, split: str = "strided", batch_scenes: int = 4, loader_threads: int = 4,
                     writer_threads: int = 2, fmt: Optional[str] = None, write: bool = True, pin_cores: bool = False,
                     pipeline=None, **pipeline_kwargs) -> Dict[str, object]:
    """The scene loop of tools/preprocess_data.py:188-297 for this rank's share of the scenes (:704-730):

        for id in my ids:  skip if `{id}.h5py` exists (:192-195)  ->  skip if not in the dataset (:197-199)
                           ->  load  ->  MVFF.fuse(..., return_obj=True) (:268)  ->  NaN rows <- query (:278-282)
                           ->  write multiview/{per_obj,obj_ids,objects_info}, pointcloud/{xyz,rgb,label,vis_mask} (:285-297)

    `source` provides `ids()`, `__contains__(id)` and `load_into(id, slot) -> {"objects_info", "queries"}`
    (SceneDirSource, or any dataset adapter that fills a PinnedSceneSlot). Loader threads fill pinned slots, the
    FusionPipeline fuses `batch_scenes` scenes per launch sequence with the copies of neighbouring batches
    overlapped, writer threads produce the files (atomic rename). Ranks never communicate on the data path; the
    returned statistics are summed over ranks with one all_reduce when torch.distributed is initialised."""
    import queue as _queue
    import threading
    import time

    from .pipeline import FusionPipeline

    rank, world = world_info()
    if pin_cores:
        pin_rank_cores()
    ids_all = list(source.ids()) if scene_ids is None else [int(i) for i in scene_ids]
    mine = rank_scene_ids(ids_all, rank, world, split)
    todo = pending_scenes(out_dir, mine)
    skipped_existing = len(mine) - len(todo)
    missing = [i for i in todo if i not in source]
    todo = [i for i in todo if i in source]
    if device is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    own_pipe = pipeline is None
    pipe = pipeline or FusionPipeline(camera_intrinsic, device=device, batch_scenes=batch_scenes, **pipeline_kwargs)
    h2d0, d2h0, launches0 = pipe.h2d_bytes, pipe.d2h_bytes, pipe.launches
    meta: Dict[int, Dict[str, object]] = {}
    errors: List[tuple] = []
    lock = threading.Lock()
    work: "_queue.Queue[Optional[int]]" = _queue.Queue()
    for i in todo:
        work.put(i)
    load_s = [0.0]

    def loader():
        while True:
            try:
                sid = work.get_nowait()
            except _queue.Empty:
                return
            slot = pipe.acquire()
            t0 = time.perf_counter()
            try:
                m = source.load_into(sid, slot)
            except Exception as exc:  # the reference skips scenes whose loading asserts (:201-205)
                pipe.release(slot)
                with lock:
                    errors.append((sid, repr(exc)))
                continue
            with lock:
                meta[sid] = m
                load_s[0] += time.perf_counter() - t0
            pipe.submit(slot, tag=sid)

    to_write: "_queue.Queue[Optional[object]]" = _queue.Queue(maxsize=4 * max(1, writer_threads))
    written = [0, 0]

    def writer():
        while True:
            r = to_write.get()
            if r is None:
                return
            m = meta.pop(r.tag)
            pts, cols, labs = r.filtered()
            path = write_scene(out_dir, r.tag, r.mv_feats_obj, m["queries"], pts, cols, labs, r.visibility_mask,
                               objects_info=m.get("objects_info"), fmt=fmt)
            with lock:
                written[0] += 1
                written[1] += os.path.getsize(path)

    t_start = time.perf_counter()
    loaders = [threading.Thread(target=loader, daemon=True) for _ in range(max(1, loader_threads))]
    writers = [threading.Thread(target=writer, daemon=True) for _ in range(max(1, writer_threads) if write else 0)]
    for t in loaders + writers:
        t.start()

    def closer():
        for t in loaders:
            t.join()
        pipe.finish()

    threading.Thread(target=closer, daemon=True).start()
    fused = 0
    for r in pipe.results():
        if r.error is not None:  # the reference raises here (IndexError, quirk q7); the driver records it and carries on
            with lock:
                errors.append((r.tag, repr(r.error)))
                meta.pop(r.tag, None)
            continue
        fused += 1
        if write:
            to_write.put(r)
    for _ in writers:
        to_write.put(None)
    for t in writers:
        t.join()
    seconds = time.perf_counter() - t_start
    if own_pipe:
        pipe.close()
    stats = {"rank": rank, "world": world, "assigned": len(mine), "skipped_existing": skipped_existing,
             "skipped_missing": len(missing), "fused": fused, "written": written[0], "written_bytes": written[1],
             "errors": errors, "seconds": seconds, "h2d_bytes": pipe.h2d_bytes - h2d0, "d2h_bytes": pipe.d2h_bytes - d2h0,
             "launches": pipe.launches - launches0, "loader_seconds": load_s[0]}
    if world > 1:
        dev = torch.device(device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([len(mine), skipped_existing, len(missing), fused, written[0], written[1], pipe.h2d_bytes - h2d0,
                          pipe.d2h_bytes - d2h0, len(errors)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tmax = torch.tensor([seconds], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        keys = ("assigned", "skipped_existing", "skipped_missing", "fused", "written", "written_bytes", "h2d_bytes",
                "d2h_bytes", "n_errors")
        stats["job"] = dict({k: int(v) for k, v in zip(keys, t.tolist())}, seconds=float(tmax.item()))
    return stats


def _main(argv=None):
    """`torchrun --nproc-per-node N -m dropclip_b200.shard --scenes DIR --out DIR` : one rank per GPU."""
    import argparse
    import json
    ap = argparse.ArgumentParser(description=_main.__doc__)
    ap.add_argument("--scenes", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--split", default="strided", choices=["strided", "reference"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--loaders", type=int, default=4)
    ap.add_argument("--writers", type=int, default=2)
    ap.add_argument("--fmt", default=None, choices=[None, "h5py", "npz"])
    args = ap.parse_args(argv)
    from .scenes import MVTOD_INTRINSIC
    if "RANK" in os.environ and not dist.is_initialized():
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    stats = run_scene_driver(SceneDirSource(args.scenes), args.out, MVTOD_INTRINSIC, split=args.split, batch_scenes=args.batch,
                             loader_threads=args.loaders, writer_threads=args.writers, fmt=args.fmt, pin_cores=True)
    if stats["rank"] == 0:
        print(json.dumps({k: v for k, v in stats.items() if k != "errors"} | {"n_errors": len(stats["errors"])}))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    _main()
