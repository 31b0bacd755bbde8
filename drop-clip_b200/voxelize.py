"""GPU replacement for the MinkowskiEngine calls the dataset makes:
`ME.utils.sparse_quantize` (data/dataset_blender.py:406-414, data/dataset.py:164-172) and
`ME.utils.sparse_collate` (data/dataset_blender.py:450-461).

Same argument names and return order as ME 0.5.x. Voxels come out in order of first occurrence
(ME's canonical order); see oracle/projections_ref.py for the parity-unpinned caveat.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .engine import FusionEngine

_engine: Optional[FusionEngine] = None


def _eng() -> FusionEngine:
    global _engine
    if _engine is None:
        _engine = FusionEngine("cuda")
    return _engine


def _to_dev(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(_eng().device).contiguous()


def sparse_quantize_batch(coordinates: Sequence, features: Optional[Sequence] = None, labels: Optional[Sequence] = None,
                          ignore_label: int = -100, quantization_size: float = 1.0):
    """Quantises a whole batch of samples in one launch sequence. Returns a list of per-sample
    tuples (coords int32 (M,3), features[unique_map] | None, labels | None, unique_map, inverse_map),
    tensors on the GPU."""
    eng = _eng()
    xyz = [_to_dev(c, torch.float32).reshape(-1, 3) for c in coordinates]
    sizes = [int(x.shape[0]) for x in xyz]
    off_host = np.zeros(len(xyz) + 1, dtype=np.int64)
    np.cumsum(sizes, out=off_host[1:])
    off = torch.from_numpy(off_host).to(eng.device)
    allxyz = torch.cat(xyz) if xyz else torch.zeros((0, 3), dtype=torch.float32, device=eng.device)
    lab = torch.cat([_to_dev(l, torch.int32).reshape(-1) for l in labels]) if labels is not None else None
    vox = eng.voxelize(allxyz, off, float(quantization_size), lab, ignore_label)
    voff = vox["voxel_off"].cpu().numpy()
    if (voff < 0).any():
        raise RuntimeError("sparse_quantize: a voxel coordinate fell outside [-2^20, 2^20)")
    feats_out = None
    if features is not None:
        allf = torch.cat([_to_dev(f).reshape(s, -1) for f, s in zip(features, sizes)])
        feats_out = eng.voxel_gather(allf, off, vox, int(voff[-1]))
    out = []
    for b in range(len(xyz)):
        s0, m = int(off_host[b]), int(voff[b + 1] - voff[b])
        out.append((vox["coords"][s0:s0 + m], None if feats_out is None else feats_out[voff[b]:voff[b + 1]],
                    None if lab is None else vox["voxel_labels"][s0:s0 + m], vox["unique_map"][s0:s0 + m],
                    vox["inverse_map"][s0:s0 + sizes[b]]))
    return out


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False, return_inverse=False,
                    return_maps_only=False, quantization_size=None, device="cuda"):
    """ME.utils.sparse_quantize for one sample. Return tuple follows ME: coords[, features][, labels]
    [, unique_map][, inverse_map]; results are CPU tensors when the input was CPU/numpy."""
    was_cuda = isinstance(coordinates, torch.Tensor) and coordinates.is_cuda
    c = torch.as_tensor(coordinates)
    if quantization_size is None:
        qs = 1.0
        if not c.is_floating_point():
            c = c.to(torch.float32)  # integer coordinates are already voxel indices
    else:
        qs = float(quantization_size)
    (coords, feats, vlab, umap, imap), = sparse_quantize_batch([c], None if features is None else [features],
                                                                None if labels is None else [labels], ignore_label, qs)
    back = (lambda t: t) if was_cuda else (lambda t: None if t is None else t.cpu())
    coords, feats, vlab, umap, imap = back(coords), back(feats), back(vlab), back(umap), back(imap)
    if return_maps_only:
        return (umap, imap) if return_inverse else umap
    res = [coords]
    if features is not None:
        res.append(feats)
    if labels is not None:
        res.append(vlab)
    if return_index:
        res.append(umap)
    if return_inverse:
        res.append(imap)
    return res[0] if len(res) == 1 else tuple(res)


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """ME.utils.sparse_collate: prepend the batch index column and concatenate."""
    cs, fs, ls = [], [], []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c).to(dtype)
        cs.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c], dim=1))
        fs.append(torch.as_tensor(feats[b]))
        if labels is not None:
            ls.append(torch.as_tensor(labels[b]))
    bc, bf = torch.cat(cs, 0), torch.cat(fs, 0)
    if device is not None:
        bc, bf = bc.to(device), bf.to(device)
    if labels is not None:
        return bc, bf, torch.cat(ls, 0)
    return bc, bf
