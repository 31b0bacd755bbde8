"""Drop-in for the reference's `models/similarity.py` (class ClipSimilarity).

Same constructor, class constants, method names, argument defaults and quirks (q15-q20): `x or
self.x` defaulting, in-place normalisation of `vis_feats`, `.squeeze()`-shaped results. The text
tower stays whatever CLIP model the caller loads (out of scope); everything after
`encode_text` - the point x prompt GEMM, paired softmax, min-max, threshold - runs in libdropclip's
tcgen05 kernel with a fused epilogue, so the (N x P) similarity matrix is never written unless
`compute_similarity` is asked for it.

Reference lines mirrored: models/similarity.py:1-101.
"""
from __future__ import annotations

import torch

from . import _lib
from .engine import FusionEngine

DEVICE = torch.device("cuda")  # the reference's expression always evaluates to 'cuda' (quirk q20)



def _load_clip(model_name, device):
    """The reference imports its vendored CLIP (`models.features.clip`); use it when importable."""
    try:
        from models.features.clip import clip  # the reference tree on sys.path
    except Exception as exc:  # pragma: no cover - depends on the caller's environment
        raise ImportError(
            "ClipSimilarity needs the CLIP text tower of the host application "
            "(`models.features.clip`); pass `model=` / `tokenize=` to use another encoder") from exc
    model, _ = clip.load(model_name, device=device)
    return model, clip.tokenize


class ClipSimilarity(object):

    NEGATIVE_PROMPT_GENERIC = ["object", "thing", "texture", "stuff"]
    SOFTMAX_TEMP = 0.1

    def __init__(self, model_name="ViT-L/14@336px", method="paired", threshold=0.7, norm_vis_feat=True, device=DEVICE,
                 model=None, tokenize=None):
        self.device = device
        self.threshold = threshold
        self.method = method
        self.norm_vis_feat = norm_vis_feat
        if model is None:
            model, tokenize = _load_clip(model_name, device)
            print(f"Loaded CLIP model {model_name}")
        self.model = model.eval().to(device)
        self.tokenize = tokenize
        self._engine = None

    # ------------------------------------------------------------------ helpers
    def _eng(self) -> FusionEngine:
        if getattr(self, "_engine", None) is None:
            self._engine = FusionEngine(self.device)
        return self._engine

    def _tok(self, text):
        tok = getattr(self, "tokenize", None)
        if tok is None:
            from models.features.clip import clip
            tok = clip.tokenize
        return tok(text)

    def _encode(self, qpos, qneg):
        """models/similarity.py:33-45: tokenise, encode, L2-normalise (tiny P x C tensors)."""
        qp = self.model.encode_text(self._tok(qpos).to(self.device))
        qp /= qp.norm(dim=-1, keepdim=True)
        if qneg is None:
            return qp, None
        assert isinstance(qneg, list), "qneg argument should be list or None"
        if not len(qneg):
            qneg = self.NEGATIVE_PROMPT_GENERIC
        qn = self.model.encode_text(self._tok(qneg).to(self.device))
        qn /= qn.norm(dim=-1, keepdim=True)
        return qp, qn

    def _check_feats(self, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise RuntimeError("dropclip_b200 grounding needs CUDA feature tensors; there is no CPU fallback")
        if x.dtype not in (torch.float16, torch.float32) or x.dim() != 2 or not x.is_contiguous():
            raise RuntimeError("vis_feats must be a contiguous (N, C) fp16 or fp32 tensor")

    # ------------------------------------------------------------------ a15
    @torch.no_grad()
    def compute_similarity(self, vis_feat_norm, qpos, qneg=None, softmax_temp=None, method="paired"):
        softmax_temp = softmax_temp or self.SOFTMAX_TEMP
        self._check_feats(vis_feat_norm)
        qp, qn = self._encode(qpos, qneg)
        eng = self._eng()
        text = qp if qn is None else torch.cat([qp, qn], dim=0)
        text = text.to(vis_feat_norm.device)
        if qn is not None and method == "paired":
            out, _, _ = eng.ground(vis_feat_norm, text, _lib.DC_GROUND_PAIRED, softmax_temp, normalize=False)
            return out.view(-1, 1).to(vis_feat_norm.dtype)
        if qn is not None and method != "argmax":
            return None  # the reference falls off the end of the if/elif chain
        out, _, _ = eng.ground(vis_feat_norm, text, _lib.DC_GROUND_RAW, softmax_temp, normalize=False)
        return out.to(vis_feat_norm.dtype)

    # ------------------------------------------------------------------ a16
    @torch.no_grad()
    def predict(self, vis_feats, qpos, qneg=None, norm_vis_feat=None, method=None, threshold=None):
        method = method or self.method
        threshold = threshold or self.threshold
        norm_vis_feat = norm_vis_feat or self.norm_vis_feat
        self._check_feats(vis_feats)
        qp, qn = self._encode(qpos, qneg)
        eng = self._eng()
        text = (qp if qn is None else torch.cat([qp, qn], dim=0)).to(vis_feats.device)
        n = vis_feats.shape[0]
        if qneg is None or (qneg is not None and method == "paired"):
            mode = _lib.DC_GROUND_RAW if qn is None else _lib.DC_GROUND_PAIRED
            sims, pred = eng.predict(vis_feats, text, mode, self.SOFTMAX_TEMP, bool(norm_vis_feat), threshold)
            pred = pred.view(torch.bool)
            if n == 1:  # .squeeze() in the reference makes these 0-d (quirk q18)
                pred, sims = pred.reshape(()), sims.reshape(())
            return pred, sims
        elif qneg is not None and method == "argmax":
            if n == 1:
                raise IndexError("too many indices for tensor of dimension 1")  # reference behaviour (q18)
            out, pred = eng.predict(vis_feats, text, _lib.DC_GROUND_ARGMAX, self.SOFTMAX_TEMP, bool(norm_vis_feat), threshold)
            return pred.view(torch.bool), out


# ---------------------------------------------------------------------- a18
@torch.no_grad()
def class_similarity(vis_feat: torch.Tensor, txt_feat: torch.Tensor, return_sims: bool = True, engine: FusionEngine = None):
    """`_get_similarity` + the arg max that follows it in the reference's evaluation loops
    (engine/distil.py:244-246,289-290, tools/validate_upper_bound.py:59-61,101-102):

        txt_feat /= txt_feat.norm(dim=-1, keepdim=True)       # IN PLACE on the class table the caller passes
        sims = vis_feat @ txt_feat.T                          # (M, K) fp32, vis_feat is NOT normalised
        pred = torch.max(sims, 1)[1]                          # (M,) int64

    One tcgen05 GEMM whose epilogue keeps each row's arg max, so the (M, K) matrix is written only when
    `return_sims` (the reference feeds it to the cross-entropy criterion, engine/distil.py:302).
    Returns (sims | None, pred)."""
    if not (isinstance(vis_feat, torch.Tensor) and vis_feat.is_cuda and isinstance(txt_feat, torch.Tensor) and txt_feat.is_cuda):
        raise RuntimeError("dropclip_b200 grounding needs CUDA tensors; there is no CPU fallback")
    if vis_feat.dim() != 2 or txt_feat.dim() != 2 or vis_feat.shape[1] != txt_feat.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(vis_feat.shape)} and {tuple(txt_feat.T.shape)})")
    txt_feat /= txt_feat.norm(dim=-1, keepdim=True)  # K x C, tiny; in place like the reference
    eng = engine or FusionEngine(vis_feat.device)
    x = vis_feat if vis_feat.dtype in (torch.float16, torch.float32) else vis_feat.float()
    x = x.contiguous()
    if x.shape[0] == 0:
        return (torch.empty((0, txt_feat.shape[0]), dtype=torch.float32, device=x.device) if return_sims else None,
                torch.empty(0, dtype=torch.int64, device=x.device))
    sims, pred, _ = eng.ground(x, txt_feat, _lib.DC_GROUND_CLASS, 0.1, normalize=False, want_matrix=return_sims)
    if sims is not None and vis_feat.dtype != torch.float32 and vis_feat.dtype == txt_feat.dtype:
        sims = sims.to(vis_feat.dtype)
    return sims, pred


def _get_similarity(vis_feat, txt_feat):
    """Name and signature of the closure in engine/distil.py:244-246 / tools/validate_upper_bound.py:59-61."""
    return class_similarity(vis_feat, txt_feat, return_sims=True)[0]
