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
