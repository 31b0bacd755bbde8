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


def contiguous_shard(start: int, end: int, rank: int, world: int) -> List[int]:
    """The reference's split of an inclusive id range into `world` contiguous chunks
    (tools/preprocess_data.py:711-726): equal chunks, the last one takes the remainder."""
    ids = list(range(start, end + 1))
    per = len(ids) // world
    lo = rank * per
    hi = (rank + 1) * per if rank < world - 1 else len(ids)
    return ids[lo:hi]


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


def output_path(out_dir: str, scene_id: int) -> str:
    return os.path.join(out_dir, "{:0>6}.npz".format(scene_id))


def pending_scenes(out_dir: str, scene_ids: Sequence[int]) -> List[int]:
    """Restart semantics of tools/preprocess_data.py:192-195: a scene whose output exists is skipped."""
    return [i for i in scene_ids if not os.path.isfile(output_path(out_dir, i))]


def write_scene(out_dir: str, scene_id: int, per_obj: np.ndarray, query: np.ndarray, xyz, rgb, label, vis_mask,
                objects_info: Optional[str] = None) -> str:
    """Same groups/keys/dtypes as the reference's h5 file (tools/preprocess_data.py:285-297), stored as
    .npz because h5py is not installed here; NaN rows (objects seen in no view, always row 0) are
    replaced by the query embedding first (:278-282)."""
    per_obj = np.array(per_obj, dtype=np.float32, copy=True)
    bad = np.isnan(per_obj).any(axis=1)
    per_obj[bad] = np.asarray(query, dtype=np.float32)[bad]
    os.makedirs(out_dir, exist_ok=True)
    path = output_path(out_dir, scene_id)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **{
        "multiview/per_obj": per_obj,
        "multiview/obj_ids": np.arange(per_obj.shape[0], dtype=np.uint8),
        "multiview/objects_info": np.asarray(objects_info or ""),
        "pointcloud/xyz": np.asarray(xyz, dtype=np.float32),
        "pointcloud/rgb": np.asarray(rgb, dtype=np.float32),
        "pointcloud/label": np.asarray(label, dtype=np.uint8),
        "pointcloud/vis_mask": np.asarray(vis_mask, dtype=np.float32),
    })
    os.replace(tmp, path)  # atomic: a killed worker never leaves a half-written file that would be skipped
    return path
