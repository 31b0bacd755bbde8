"""Scene pipeline: the throughput form of the object-level fusion loop (tools/preprocess_data.py:188-297).

The reference fuses one scene per Python iteration from freshly loaded numpy arrays. Per scene that is 275 MB of
input (73 fp32 depth maps + 73 int64 instance maps) for 0.06 ms of GPU work, so the loop runs at the speed the bytes
reach the GPU. Here the loader writes each scene ONCE, straight into library-owned pinned memory (`PinnedSceneSlot`:
fp32 depths, instance ids as uint8 - the 1-byte form `SceneBatch.from_host` narrows to as well, int64 only when an id
does not fit), and `FusionPipeline` moves slots through three stages on three CUDA streams:

    H2D (copy stream)  ->  fuse a BATCH of scenes per launch sequence (compute stream)  ->  D2H of the results (copy-out stream)

with three device arenas (inputs + pinned result buffers) in flight, so the PCIe copies of batch k+1 and k-1 overlap the
kernels of batch k and nothing is allocated in steady state. A dispatcher thread issues the work, a completion
thread waits on the events and hands out `SceneResult`s; slots go back to the free list as soon as their H2D copies
have finished. Results are exactly what `MultiviewFeatureFusion.fuse(..., return_obj=True)` returns for the scene
(the visibility mask as uint8 - the reference's int64 values 0/1 -, points/colors/labels filtered lazily).

No CPU fallback: every stage is libdropclip kernels + async copies; a missing library or device raises.
"""
from __future__ import annotations

import ctypes
import queue
import threading
from dataclasses import dataclass, field
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import FusionEngine, PinnedStaging, SceneBatch, _labels_as_int64, intrinsic_matrix

__all__ = ["PinnedSceneSlot", "SceneResult", "FusionPipeline"]


def _pinned(shape, dtype, pin: bool = True) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, pin_memory=pin)


class PinnedSceneSlot:
    """Pinned host buffers for one scene at fixed capacities. The loader fills them in place (`depths[v][...] = ...`,
    or `fill(...)` from the reference's containers) and hands the slot to `FusionPipeline.submit`."""

    def __init__(self, height: int, width: int, max_views: int, max_points: int, max_rows: int, max_queries: int,
                 feat_dim: int = 768, feat_dtype=torch.float16, pinned: bool = True):
        """`pinned=False` gives ordinary host memory (host-logic tests of loaders without a GPU; the pipeline itself
        always allocates pinned slots)."""
        self.height, self.width = height, width
        self.cap = dict(views=max_views, points=max_points, rows=max_rows, queries=max_queries)
        self._pin = pinned
        self.t_depths = _pinned((max_views, height, width), torch.float32, pinned)
        self.t_segs = _pinned((max_views, height, width), torch.uint8, pinned)
        self.t_segs_wide: Optional[torch.Tensor] = None  # int64 maps, allocated on first use (an id outside [0, 255])
        self.t_inv_poses = _pinned((max_views, 16), torch.float64, pinned)
        self.t_points = _pinned((max_points, 3), torch.float64, pinned)
        self.t_labels = _pinned((max_points,), torch.int64, pinned)
        self.t_feats = _pinned((max_rows, feat_dim), feat_dtype, pinned)
        self.t_queries = _pinned((max_queries, feat_dim), torch.float32, pinned)
        # numpy views for loaders
        self.depths, self.segs = self.t_depths.numpy(), self.t_segs.numpy()
        self.inv_poses, self.points, self.labels = self.t_inv_poses.numpy(), self.t_points.numpy(), self.t_labels.numpy()
        self.feats, self.queries = self.t_feats.numpy(), self.t_queries.numpy()
        self.feat_rows = np.zeros(max_views, dtype=np.int64)
        self.n_views = self.n_points = self.n_queries = 0
        self.wide_segs = False
        self.colors: Any = None       # caller's arrays kept by reference for the lazily filtered outputs
        self.points_src: Any = None
        self.labels_src: Any = None
        self.tag: Any = None

    # ------------------------------------------------------------------ filling
    def set_poses(self, camera_poses: Sequence[np.ndarray]) -> None:
        """camera->world matrices; inverted in their own dtype like utils/transforms.py:54, stored as fp64 (exact)."""
        poses = [np.asarray(p) for p in camera_poses]
        inv = np.linalg.inv(np.stack(poses)) if all(p.dtype == poses[0].dtype for p in poses) else \
            np.stack([np.linalg.inv(p) for p in poses])
        self.inv_poses[:len(poses)] = inv.reshape(len(poses), 16)

    def fill(self, points, colors, labels, depths, seg_masks, camera_poses, mv_features, query_embeddings, threads: int = 0):
        """Copies one scene given in the reference's containers (the argument list of fuse()) into the slot.
        Loaders that decode files should rather write into `depths` / `segs` / ... directly."""
        lib = _lib.load()
        V, N = len(depths), int(np.shape(points)[0])
        Q = int(query_embeddings.shape[0])
        rows = [int(f.shape[0]) for f in mv_features]
        if V > self.cap["views"] or N > self.cap["points"] or Q > self.cap["queries"] or sum(rows) > self.cap["rows"]:
            raise ValueError(f"scene (V={V}, N={N}, Q={Q}, rows={sum(rows)}) exceeds the slot capacity {self.cap}")
        threads = threads or PinnedStaging._host_threads()
        hw = self.height * self.width
        dl = [np.ascontiguousarray(d, dtype=np.float32) for d in depths]
        srcs = (ctypes.c_void_p * V)(*[d.ctypes.data for d in dl])
        _lib.check(lib.dc_host_gather_copy(srcs, V, hw * 4, ctypes.c_void_p(self.t_depths.data_ptr()), threads))
        self.store_segs(seg_masks, threads)
        self.set_poses(camera_poses)
        r0 = 0
        for v, f in enumerate(mv_features):
            if f.shape[-1] != self.t_feats.shape[1]:
                raise RuntimeError(f"The expanded size of the tensor ({self.t_feats.shape[1]}) must match the existing size ({f.shape[-1]})")
            self.t_feats[r0:r0 + rows[v]].copy_(f)
            r0 += rows[v]
        self.t_queries[:Q].copy_(query_embeddings)
        return self.set_scene(V, N, Q, rows, points, colors, labels)

    def store_segs(self, seg_masks, threads: int = 0) -> None:
        """Instance maps (a list of (H,W) arrays or one (V,H,W) array, any integer dtype) into the slot: as uint8 when
        every id fits a byte, else unchanged as int64 (`wide_segs`; e.g. a -1 background, which np.unique()[1:] drops)."""
        lib = _lib.load()
        threads = threads or PinnedStaging._host_threads()
        hw = self.height * self.width
        sl = [np.ascontiguousarray(s.cpu().numpy() if isinstance(s, torch.Tensor) else s) for s in seg_masks]
        V = len(sl)
        if V > self.cap["views"]:
            raise ValueError(f"{V} views exceed the slot capacity {self.cap['views']}")
        self.wide_segs = False
        if all(s.dtype == np.uint8 for s in sl):
            for v, s in enumerate(sl):
                self.segs[v] = s
            return
        sl = [s.astype(np.int64, copy=False) for s in sl]
        srcs = (ctypes.c_void_p * V)(*[s.ctypes.data for s in sl])
        bad = ctypes.c_int(0)
        _lib.check(lib.dc_host_gather_narrow_i64_u8(srcs, V, hw, ctypes.c_void_p(self.t_segs.data_ptr()), threads,
                                                    ctypes.byref(bad)))
        if bad.value:  # an id outside [0, 255]: ship the int64 maps unchanged
            if self.t_segs_wide is None:
                self.t_segs_wide = _pinned((self.cap["views"], self.height, self.width), torch.int64, self._pin)
            _lib.check(lib.dc_host_gather_copy(srcs, V, hw * 8, ctypes.c_void_p(self.t_segs_wide.data_ptr()), threads))
            self.wide_segs = True

    def set_scene(self, n_views: int, n_points: int, n_queries: int, feat_rows, points, colors, labels):
        """Extents + the per-point arrays, for loaders that wrote depths / segs / feats / queries / poses in place.
        `points`, `colors`, `labels` are kept by reference for the lazily filtered outputs (they must not alias the slot)."""
        rows = [int(r) for r in feat_rows]
        V, N, Q = int(n_views), int(n_points), int(n_queries)
        if V > self.cap["views"] or N > self.cap["points"] or Q > self.cap["queries"] or sum(rows) > self.cap["rows"] or len(rows) != V:
            raise ValueError(f"scene (V={V}, N={N}, Q={Q}, rows={sum(rows)}) exceeds the slot capacity {self.cap}")
        self.points[:N] = np.asarray(points, dtype=np.float64).reshape(N, 3)
        self.labels[:N] = _labels_as_int64(labels)
        self.feat_rows[:V] = rows
        self.n_views, self.n_points, self.n_queries = V, N, Q
        self.points_src, self.colors, self.labels_src = points, colors, labels
        return self

    def input_bytes(self) -> int:
        """Bytes the H2D stage moves for this scene."""
        hw = self.height * self.width
        rows = int(self.feat_rows[:self.n_views].sum())
        return (self.n_views * hw * (4 + (8 if self.wide_segs else 1)) + self.n_views * 128 + self.n_points * (24 + 8) +
                rows * self.t_feats.shape[1] * self.t_feats.element_size() + self.n_queries * self.t_feats.shape[1] * 4)


@dataclass
class SceneResult:
    """What fuse(..., return_obj=True) returns for one scene, on the host (pinned buffers recycled after `ttl` further
    results have been handed out - copy what must live longer)."""
    tag: Any
    mv_feats_obj: np.ndarray      # (Q, C) fp32, NaN rows for objects seen in no view (quirk q10)
    weight_obj: np.ndarray        # (Q, V) fp32
    visibility_mask: np.ndarray   # (V, N') uint8 (the reference's int64 0/1 values)
    keep: np.ndarray              # (N,) bool: points seen in at least one view
    error: Optional[Exception] = None
    _src: Tuple[Any, Any, Any] = field(default=(None, None, None), repr=False)

    def filtered(self):
        """(points[keep], colors[keep], labels[keep]) like the second tuple fuse() returns."""
        return tuple(None if a is None else np.asarray(a)[self.keep] for a in self._src)


class _Arena:
    """Device buffers for one batch in flight + the pinned buffers its results are copied into."""

    def __init__(self, dev, B, slot0: PinnedSceneSlot):
        c, H, W = slot0.cap, slot0.height, slot0.width
        dim, fdt = slot0.t_feats.shape[1], slot0.t_feats.dtype
        self.depths = torch.empty((B * c["views"], H, W), dtype=torch.float32, device=dev)
        self.segs = torch.empty((B * c["views"], H, W), dtype=torch.uint8, device=dev)
        self.segs_wide: Optional[torch.Tensor] = None
        self.inv_poses = torch.empty((B * c["views"], 16), dtype=torch.float64, device=dev)
        self.points = torch.empty((B * c["points"], 3), dtype=torch.float64, device=dev)
        self.labels = torch.empty((B * c["points"],), dtype=torch.int64, device=dev)
        self.feats = torch.empty((B * c["rows"], dim), dtype=fdt, device=dev)
        self.queries = torch.empty((B * c["queries"], dim), dtype=torch.float32, device=dev)
        self.intrinsics = torch.empty((B, 9), dtype=torch.float64, device=dev)
        # packed offset arrays: five (B + 1) int64 prefixes, feat_off (B V + 1) int64, view_scene (B V) int32, 16-byte slots
        self.meta = torch.empty(5 * (8 * (B + 1) + 16) + 8 * (B * c["views"] + 1) + 16 + 4 * B * c["views"] + 16,
                                dtype=torch.uint8, device=dev)
        self.h_meta = _pinned(self.meta.shape, torch.uint8)
        self.h_K = _pinned((B, 9), torch.float64)
        # results (upper bounds)
        self.h_mask = _pinned((B * c["views"] * c["points"],), torch.uint8)
        self.h_any = _pinned((B * c["points"],), torch.uint8)
        self.h_fused = _pinned((B * c["queries"], dim), torch.float32)
        self.h_weight = _pinned((B * c["queries"] * c["views"],), torch.float32)
        self.h_status = _pinned((B * c["views"],), torch.int32)
        self.h_off = _pinned((2, B + 1), torch.int64)  # kept_off, out_off
        self.copied_in = torch.cuda.Event()
        self.computed = torch.cuda.Event()
        self.copied_out = torch.cuda.Event()
        self.free = threading.Event()
        self.free.set()


class FusionPipeline:
    """See the module docstring. Typical use (what shard.run_scene_driver does):

        pipe = FusionPipeline(intrinsic, device="cuda:0", batch_scenes=4)
        for scene in loader:                      # any thread(s)
            slot = pipe.acquire()                 # blocks while all slots are in flight
            slot.fill(*scene.fuse_arguments)      # or decode straight into slot.depths / slot.segs / ...
            pipe.submit(slot, tag=scene.id)
        pipe.finish()
        for res in pipe.results():                # SceneResult per scene, in submission order
            ...
    """

    def __init__(self, camera_intrinsic: Dict[str, float], device="cuda", image_size=(480, 640), batch_scenes: int = 4,
                 n_slots: Optional[int] = None, max_views: int = 73, max_points: int = 100_000, max_rows: Optional[int] = None,
                 max_queries: int = 21, feat_dim: int = 768, feat_dtype=torch.float16, visibility_threshold: float = 0.05,
                 use_visibility: bool = False, use_similarity: bool = True, use_sim_kernel: Optional[str] = "max",
                 n_arenas: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("dropclip_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError(f"dropclip_b200 runs on CUDA devices only (got device={device!r}); there is no CPU fallback")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.H, self.W = int(image_size[0]), int(image_size[1])
        self.K = intrinsic_matrix(camera_intrinsic).reshape(9)
        self.B = int(batch_scenes)
        self.threshold, self.use_visibility, self.use_similarity = visibility_threshold, use_visibility, use_similarity
        self.sim_kernel = use_sim_kernel if use_similarity else None
        if use_similarity and use_sim_kernel not in ("max", "mean"):
            raise ValueError("Please set method in [mean, max]")
        max_rows = max_rows or max_views * max(max_queries - 1, 1)
        n_slots = n_slots or 3 * self.B
        with torch.cuda.device(self.dev):
            self.eng = FusionEngine(self.dev)
            self._slots = [PinnedSceneSlot(self.H, self.W, max_views, max_points, max_rows, max_queries, feat_dim, feat_dtype)
                           for _ in range(n_slots)]
            self._arenas = [_Arena(self.dev, self.B, self._slots[0]) for _ in range(n_arenas)]
            self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self._free: "queue.Queue[PinnedSceneSlot]" = queue.Queue()
        for s in self._slots:
            self._free.put(s)
        self._submitted: "queue.Queue[Optional[PinnedSceneSlot]]" = queue.Queue()
        self._inflight: "queue.Queue[Optional[tuple]]" = queue.Queue()
        # bounded: a slow consumer (e.g. the file writers of shard.run_scene_driver) holds back the completion thread,
        # hence the arenas, hence the slots - instead of results piling up in host memory
        self._results: "queue.Queue[Optional[SceneResult]]" = queue.Queue(maxsize=8 * self.B + 32)
        self.h2d_bytes = self.d2h_bytes = 0
        self.launches = 0
        self._error: Optional[BaseException] = None
        self._threads = [threading.Thread(target=self._dispatch, daemon=True), threading.Thread(target=self._complete, daemon=True)]
        for t in self._threads:
            t.start()

    # ------------------------------------------------------------------ public
    def acquire(self) -> PinnedSceneSlot:
        while True:
            try:
                return self._free.get(timeout=0.2)
            except queue.Empty:
                if self._error is not None:  # a stage died: do not wait for slots that will never come back
                    raise RuntimeError("FusionPipeline stopped") from self._error

    def release(self, slot: PinnedSceneSlot) -> None:
        """Hands an acquired slot back unused (the loader gave up on its scene)."""
        self._free.put(slot)

    def submit(self, slot: PinnedSceneSlot, tag: Any = None) -> None:
        slot.tag = tag
        self._submitted.put(slot)

    def finish(self) -> None:
        """No more scenes: flushes the partial batch; results() ends after the last scene."""
        self._submitted.put(None)

    def results(self) -> Iterator[SceneResult]:
        while True:
            r = self._results.get()
            if r is None:
                if self._error is not None:
                    raise self._error
                return
            yield r

    def close(self) -> None:
        for t in self._threads:
            t.join(timeout=5.0)

    # ------------------------------------------------------------------ stage 1+2: H2D and kernels (dispatcher thread)
    def _dispatch(self):
        try:
            torch.cuda.set_device(self.dev)
            k = 0
            done = False
            while not done:
                batch: List[PinnedSceneSlot] = []
                while len(batch) < self.B:
                    # take what is there; wait only for the first scene of a batch (a partial batch beats an idle GPU)
                    try:
                        s = self._submitted.get(block=(len(batch) == 0), timeout=None if len(batch) == 0 else 0.0005)
                    except queue.Empty:
                        break
                    if s is None:
                        done = True
                        break
                    batch.append(s)
                if batch:
                    self._issue(self._arenas[k % len(self._arenas)], batch)
                    k += 1
            self._inflight.put(None)
        except BaseException as exc:  # surface in results()
            self._error = exc
            self._inflight.put(None)

    def _issue(self, ar: _Arena, slots: List[PinnedSceneSlot]):
        ar.free.wait()
        ar.free.clear()
        n_points = [s.n_points for s in slots]
        n_views = [s.n_views for s in slots]
        n_queries = [s.n_queries for s in slots]
        feat_rows = [int(r) for s in slots for r in s.feat_rows[:s.n_views]]
        host = SceneBatch.offsets_for(n_points, n_views, n_queries, feat_rows)
        wide = any(s.wide_segs for s in slots)
        if wide and ar.segs_wide is None:
            ar.segs_wide = torch.empty(ar.segs.shape, dtype=torch.int64, device=self.dev)
        segs_dev = ar.segs_wide if wide else ar.segs
        pv, pp, pq, pr = host["view"], host["point"], host["query"], np.concatenate([[0], np.cumsum([sum(s.feat_rows[:s.n_views]) for s in slots])]).astype(np.int64)
        with torch.cuda.stream(self.s_in):
            for i, s in enumerate(slots):
                V, N, Q, R = s.n_views, s.n_points, s.n_queries, int(pr[i + 1] - pr[i])
                ar.depths[pv[i]:pv[i] + V].copy_(s.t_depths[:V], non_blocking=True)
                if wide:
                    src = s.t_segs_wide[:V] if s.wide_segs else s.t_segs[:V].to(torch.int64).pin_memory()
                    segs_dev[pv[i]:pv[i] + V].copy_(src, non_blocking=True)
                else:
                    segs_dev[pv[i]:pv[i] + V].copy_(s.t_segs[:V], non_blocking=True)
                ar.inv_poses[pv[i]:pv[i] + V].copy_(s.t_inv_poses[:V], non_blocking=True)
                ar.points[pp[i]:pp[i] + N].copy_(s.t_points[:N], non_blocking=True)
                ar.labels[pp[i]:pp[i] + N].copy_(s.t_labels[:N], non_blocking=True)
                ar.feats[pr[i]:pr[i] + R].copy_(s.t_feats[:R], non_blocking=True)
                ar.queries[pq[i]:pq[i] + Q].copy_(s.t_queries[:Q], non_blocking=True)
                self.h2d_bytes += s.input_bytes()
            # the seven offset arrays in one copy
            keys = list(host)
            offs, total = [], 0
            hm = ar.h_meta.numpy()
            for kname in keys:
                a = np.ascontiguousarray(host[kname])
                offs.append((total, a))
                hm[total:total + a.nbytes] = a.reshape(-1).view(np.uint8)
                total += (a.nbytes + 15) // 16 * 16
            ar.meta[:total].copy_(ar.h_meta[:total], non_blocking=True)
            ar.h_K.numpy()[:len(slots)] = self.K
            ar.intrinsics[:len(slots)].copy_(ar.h_K[:len(slots)], non_blocking=True)
            ar.copied_in.record(self.s_in)
        off = {}
        for kname, (o, a) in zip(keys, offs):
            tdt = torch.int32 if a.dtype == np.int32 else torch.int64
            off[kname] = ar.meta[o:o + a.nbytes].view(tdt)
        b = SceneBatch(device=self.dev, height=self.H, width=self.W, n_scenes=len(slots), n_points=n_points, n_views=n_views,
                       n_queries=n_queries, feat_rows=feat_rows, points=ar.points[:pp[-1]], depths=ar.depths[:pv[-1]],
                       inv_poses=ar.inv_poses[:pv[-1]], intrinsics=ar.intrinsics[:len(slots)], segs=segs_dev[:pv[-1]],
                       labels=ar.labels[:pp[-1]], feats=ar.feats[:pr[-1]], queries=ar.queries[:pq[-1]])
        b.off, b.off_host = off, host
        eng = self.eng
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(ar.copied_in)
            l0 = eng.launches
            res = eng.fuse_object_level(b, self.threshold, self.use_visibility, self.use_similarity, self.sim_kernel, torch.uint8,
                                        join=False)
            _, kept_off, _, out_off, cmask, _ = eng.compact_visibility(b, res["any_visible"], res["records"], res["rank"],
                                                                       torch.uint8, host_sizes=False)
            res["join"]()
            self.launches += eng.launches - l0
            ar.computed.record(self.s_compute)
        tm, tp, tq, tw, tv = int(host["mask"][-1]), int(pp[-1]), int(pq[-1]), int(host["wobj"][-1]), int(pv[-1])
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ar.computed)
            ar.h_off[0, :len(slots) + 1].copy_(kept_off, non_blocking=True)
            ar.h_off[1, :len(slots) + 1].copy_(out_off, non_blocking=True)
            ar.h_mask[:tm].copy_(cmask[:tm], non_blocking=True)  # upper bound: the kept columns are a prefix per scene block
            ar.h_any[:tp].copy_(res["any_visible"][:tp], non_blocking=True)
            ar.h_fused[:tq].copy_(res["fused"][:tq], non_blocking=True)
            ar.h_weight[:tw].copy_(res["weight_obj"][:tw], non_blocking=True)
            ar.h_status[:tv].copy_(res["view_status"][:tv], non_blocking=True)
            ar.copied_out.record(self.s_out)
            for t in (cmask, res["any_visible"], res["fused"], res["weight_obj"], res["view_status"], kept_off, out_off,
                      res["records"], res["rank"]):
                t.record_stream(self.s_out)
        self.d2h_bytes += tm + tp + tq * ar.h_fused.shape[1] * 4 + tw * 4 + tv * 4 + 16 * (len(slots) + 1)
        self._inflight.put((ar, slots, host, res))

    # ------------------------------------------------------------------ stage 3: completion thread
    def _complete(self):
        try:
            torch.cuda.set_device(self.dev)
            while True:
                item = self._inflight.get()
                if item is None:
                    break
                ar, slots, host, res = item
                ar.copied_in.synchronize()
                srcs = [(s.tag, s.points_src, s.colors, s.labels_src, s.n_views, s.n_points, s.n_queries,
                         [int(r) for r in s.feat_rows[:s.n_views]]) for s in slots]
                for s in slots:  # inputs are on the device: the loader may refill these slots now
                    s.points_src = s.colors = s.labels_src = None
                    self._free.put(s)
                ar.copied_out.synchronize()
                kept = ar.h_off[0].numpy()
                out_off = ar.h_off[1].numpy()
                pp, pq, pw, pv = host["point"], host["query"], host["wobj"], host["view"]
                dim = ar.h_fused.shape[1]
                for i, (tag, pts, cols, labs, V, N, Q, rows) in enumerate(srcs):
                    n_kept = int(kept[i + 1] - kept[i])
                    status = ar.h_status.numpy()[pv[i]:pv[i] + V]
                    err = None
                    if (status & 1).any():
                        err = IndexError(f"index out of bounds: view {int(np.flatnonzero(status & 1)[0])} contains an instance id outside [0, {Q})")
                    elif (status & 2).any():
                        v = int(np.flatnonzero(status & 2)[0])
                        err = IndexError(f"index {rows[v]} is out of bounds for dimension 0 with size {rows[v]}")
                    o0 = int(out_off[i])
                    self._results.put(SceneResult(
                        tag=tag,
                        mv_feats_obj=ar.h_fused.numpy()[pq[i]:pq[i] + Q].copy(),
                        weight_obj=ar.h_weight.numpy()[pw[i]:pw[i] + Q * V].reshape(Q, V).copy(),
                        visibility_mask=ar.h_mask.numpy()[o0:o0 + V * n_kept].reshape(V, n_kept).copy(),
                        keep=ar.h_any.numpy()[pp[i]:pp[i] + N].astype(bool),
                        error=err, _src=(pts, cols, labs)))
                del res
                ar.free.set()
            self._results.put(None)
        except BaseException as exc:
            self._error = exc
            self._results.put(None)
