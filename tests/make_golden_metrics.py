"""Writes tests/golden/metrics.npz: seeded inputs and the outputs of the UNMODIFIED reference's
`trainMetricPC` / `intersectionAndUnionGPU` (utils/misc.py). Run in the build container only
(`python tests/make_golden_metrics.py`); the reference tree does not travel to the GPU box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def metric_cases():
    """(name, preds, gts, threshold, sigmoid): lists of per-instance tensors."""
    rng = np.random.default_rng(11)
    cases = []
    sizes = [1, 7, 100, 1000, 4097, 33]
    scores = [torch.from_numpy(rng.random(n, dtype=np.float32)) for n in sizes]
    gts = [torch.from_numpy((rng.random(n) < 0.4)) for n in sizes]
    cases.append(("float_bool", scores, gts, 0.35, False))
    cases.append(("float_sigmoid", [(s - 0.5) * 6 for s in scores], [g.to(torch.int64) for g in gts], 0.5, True))
    bools = [s > 0.6 for s in scores]  # what ClipSimilarity.predict hands over (engine/distil.py:447-460)
    cases.append(("bool_int64", bools, [g.to(torch.int64) for g in gts], 0.35, False))
    weird = [s.clone() for s in scores]
    weird[2][::9] = float("nan")
    weird[3][::5] = 0.35  # exactly on the threshold
    weird[4][:] = 0.0     # empty prediction -> iou 0 unless gt empty too
    g2 = [g.clone() for g in gts]
    g2[4][:] = False      # union 0 -> 0 / 1e-6
    cases.append(("nan_edge", [w.view(-1, 1) for w in weird], [g.to(torch.float32) for g in g2], 0.35, False))
    return cases


def class_cases():
    rng = np.random.default_rng(12)
    out = []
    for name, n, k, dt in (("k44", 20000, 44, np.int64), ("k5_i32", 999, 5, np.int32), ("k44_2d", 6000, 44, np.int64)):
        pred = rng.integers(0, k, size=n).astype(dt)
        tgt = rng.integers(0, k, size=n).astype(dt)
        tgt[rng.random(n) < 0.1] = 255
        agree = rng.random(n) < 0.5
        pred[agree] = np.where(tgt[agree] == 255, pred[agree], tgt[agree])
        if name.endswith("2d"):
            pred, tgt = pred.reshape(60, 100), tgt.reshape(60, 100)
        out.append((name, pred, tgt, k))
    return out


def main():
    ref_shim.load()
    import utils.misc as misc
    g = {}
    for name, preds, gts, thr, sig in metric_cases():
        g[f"m_{name}_sizes"] = np.array([p.numel() for p in preds])
        g[f"m_{name}_pred"] = np.concatenate([p.reshape(-1).to(torch.float32).numpy() for p in preds])
        g[f"m_{name}_gt"] = np.concatenate([g_.reshape(-1).to(torch.float32).numpy() for g_ in gts])
        p_in = [p.clone() for p in preds]
        iou, precs = misc.trainMetricPC(p_in, [x.clone() for x in gts], threshold=thr, sigmoid=sig)
        g[f"m_{name}_out"] = np.array([float(iou)] + [float(x) for x in precs], dtype=np.float64)
        g[f"m_{name}_pred_after"] = np.concatenate([p.reshape(-1).to(torch.float32).numpy() for p in p_in])
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference ends with .cuda(); identity on this CPU box
    try:
        for name, pred, tgt, k in class_cases():
            p = torch.from_numpy(pred.copy())
            t = torch.from_numpy(tgt.copy())
            ai, au, at = misc.intersectionAndUnionGPU(p, t, k, 255)
            g[f"c_{name}_pred"], g[f"c_{name}_tgt"] = pred, tgt
            g[f"c_{name}_out"] = np.stack([ai.numpy(), au.numpy(), at.numpy()])
            g[f"c_{name}_pred_after"] = p.numpy()
    finally:
        torch.Tensor.cuda = cuda
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **g)
    print("wrote metrics", len(g))


if __name__ == "__main__":
    main()
