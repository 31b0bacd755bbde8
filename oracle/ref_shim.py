"""ORACLE (test infrastructure): imports the UNMODIFIED reference from /root/reference.

Only usable in the build container (the reference tree does not travel to the GPU box); used
by `tests/make_golden.py` to produce the committed golden vectors and by the `needs_reference`
tests that re-validate the restatements. Follows the recipe verified in SURVEY.md §8c: empty
module stubs for open3d / trimesh / ftfy, a table-lookup CLIP text tower and a crc32 tokenizer.
"""
from __future__ import annotations

import os
import sys
import types
import warnings
import zlib

import torch

REFERENCE_ROOT = os.environ.get("DROPCLIP_REF", "/root/reference")
VOCAB = 4096


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "feature_fusion.py"))


def load():
    """Returns the reference modules (feature_fusion, projections, similarity, transforms)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("open3d", "trimesh", "ftfy"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import utils.feature_fusion as ff  # noqa
        import utils.projections as pj  # noqa
        import utils.transforms as tf  # noqa
        import models.similarity as ms  # noqa
    return ff, pj, ms, tf


def fake_tokenize(texts):
    if isinstance(texts, str):
        texts = [texts]
    return torch.tensor([zlib.crc32(t.encode()) % VOCAB for t in texts], dtype=torch.long)


class FakeTextTower:
    """encode_text(tokens) -> table[tokens]; the table dtype must equal the feature dtype."""

    def __init__(self, dim=768, dtype=torch.float32, seed=7):
        import numpy as np
        # numpy's Generator stream is stable across machines and versions (torch's is not promised to be)
        table = np.random.default_rng(seed).standard_normal((VOCAB, dim)).astype(np.float32)
        self.table = torch.from_numpy(table).to(dtype)

    def encode_text(self, tokens):
        return self.table[tokens.cpu()].clone()

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self


def make_reference_similarity(dim=768, dtype=torch.float32, method="paired", threshold=0.7, seed=7):
    """An instance of the reference's own ClipSimilarity with the text tower replaced."""
    _, _, ms, _ = load()
    ms.clip.tokenize = fake_tokenize
    cs = object.__new__(ms.ClipSimilarity)
    cs.device, cs.threshold, cs.method, cs.norm_vis_feat = "cpu", threshold, method, True
    cs.model = FakeTextTower(dim, dtype, seed)
    return cs
