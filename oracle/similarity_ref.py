"""ORACLE (test infrastructure, never the product path): CPU restatement of DROP-CLIP's
text-prompt grounding arithmetic, i.e. everything in `models/similarity.py` that happens
after the CLIP text tower has produced prompt embeddings.

Parity status: PINNED against the unmodified reference class (run with a table-lookup text
encoder, see `oracle/ref_shim.py`) by `tests/make_golden.py` -> `tests/golden/ground_*.npz`.

Reference lines followed (under /root/reference):
  unit_rows_           models/similarity.py:35,45,77   (in-place row L2 normalisation)
  raw_similarity       models/similarity.py:48-49,67
  paired_softmax       models/similarity.py:51-61
  predict_from_embeds  models/similarity.py:70-101
  class_similarity     engine/distil.py:244-246,289-290
"""
from __future__ import annotations

from typing import Optional

import torch

SOFTMAX_TEMP = 0.1


def unit_rows_(x: torch.Tensor) -> torch.Tensor:
    x /= x.norm(dim=-1, keepdim=True)
    return x


def raw_similarity(vis: torch.Tensor, qpos: torch.Tensor, qneg: Optional[torch.Tensor]) -> torch.Tensor:
    if qneg is None:
        return vis @ qpos.T
    return vis @ torch.cat([qpos, qneg], dim=0).T


def paired_softmax(raw: torch.Tensor, temp: float = SOFTMAX_TEMP) -> torch.Tensor:
    pos, neg = raw[..., :1], raw[..., 1:]
    pairs = torch.cat([pos.broadcast_to(neg.shape), neg], dim=-1)
    prob = (pairs / temp).softmax(dim=-1)[..., :1]
    torch.nan_to_num_(prob, nan=0.0)
    out, _ = prob.min(dim=-1, keepdim=True)
    return out


def paired_closed_form(raw: torch.Tensor, temp: float = SOFTMAX_TEMP) -> torch.Tensor:
    """Same quantity without the concatenation: 1 / (Nneg + sum_j exp((neg_j - pos)/T)).
    This is the formula the CUDA epilogue evaluates; kept here so the tests can show that the
    two agree to rounding (SURVEY.md §8c measured 5.4e-7)."""
    pos, neg = raw[..., :1].float(), raw[..., 1:].float()
    s = torch.exp((neg - pos) / temp).sum(-1, keepdim=True)
    out = 1.0 / (neg.shape[-1] + s)
    return torch.nan_to_num(out, nan=0.0)


def predict_from_embeds(vis: torch.Tensor, qpos: torch.Tensor, qneg: Optional[torch.Tensor],
                        method: str = "paired", threshold: float = 0.7, norm_vis_feat: bool = True):
    """`ClipSimilarity.predict` with the text tower factored out: `qpos` (1,C) and `qneg`
    (Nneg,C) are the *un-normalised* prompt embeddings in the feature dtype."""
    if norm_vis_feat:
        unit_rows_(vis)
    qpos = unit_rows_(qpos.clone())
    if qneg is not None:
        qneg = unit_rows_(qneg.clone())
    raw = raw_similarity(vis, qpos, qneg)
    if qneg is not None and method == "paired":
        sims = paired_softmax(raw).squeeze()
    else:
        sims = raw.squeeze()
    if qneg is None or method == "paired":
        if sims.max() != sims.min():
            norm = (sims - sims.min()) / (sims.max() - sims.min())
        else:
            norm = sims / sims.max()
        return norm > threshold, norm.float()
    dif = sims[:, 0] - sims[:, 1:].mean(-1)
    if sims.max() != sims.min():
        norm = (dif - dif.min()) / (dif.max() - dif.min())
    else:
        norm = dif / dif.max()
    return torch.max(sims, 1)[1] == 0, norm.float()


def class_similarity(points_feat: torch.Tensor, class_table: torch.Tensor):
    """engine/distil.py:244-246 + argmax :290. `class_table` is normalised in place."""
    class_table /= class_table.norm(dim=-1, keepdim=True)
    sims = points_feat @ class_table.T
    return sims, sims.argmax(-1)
