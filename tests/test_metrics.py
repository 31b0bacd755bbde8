"""Metric reductions that follow grounding: oracle vs golden vectors of the unmodified
reference (CPU), CUDA drop-in vs both (GPU). Integer counts are exact; means are fp32 (rel 1e-5)."""
import numpy as np
import pytest
import torch

from tests import golden_io as gio
from tests.make_golden_metrics import class_cases, metric_cases


def _split(flat, sizes, shapes_like, dtype):
    off = np.r_[0, np.cumsum(sizes)]
    return [torch.from_numpy(flat[off[i]:off[i + 1]].copy()).to(dtype).view(shapes_like[i].shape) for i in range(len(sizes))]


def _metric_inputs(name):
    for n, preds, gts, thr, sig in metric_cases():
        if n == name:
            return preds, gts, thr, sig
    raise KeyError(name)


METRIC = ["float_bool", "float_sigmoid", "bool_int64", "nan_edge"]
CLASS = ["k44", "k5_i32", "k44_2d"]


@pytest.mark.parametrize("name", METRIC)
def test_oracle_train_metric_pc_vs_reference_golden(name):
    from oracle import metrics_ref
    z = gio.load("metrics.npz")
    preds, gts, thr, sig = _metric_inputs(name)
    assert np.array_equal(np.concatenate([p.reshape(-1).float().numpy() for p in preds]), z[f"m_{name}_pred"], equal_nan=True)
    p_in = [p.clone() for p in preds]
    iou, precs = metrics_ref.train_metric_pc(p_in, [g.clone() for g in gts], threshold=thr, sigmoid=sig)
    got = np.array([float(iou)] + [float(x) for x in precs])
    assert np.array_equal(got, z[f"m_{name}_out"])
    assert np.array_equal(np.concatenate([p.reshape(-1).float().numpy() for p in p_in]), z[f"m_{name}_pred_after"], equal_nan=True)


@pytest.mark.parametrize("name", CLASS)
def test_oracle_intersection_and_union_vs_reference_golden(name):
    from oracle import metrics_ref
    z = gio.load("metrics.npz")
    p = torch.from_numpy(z[f"c_{name}_pred"].copy())
    t = torch.from_numpy(z[f"c_{name}_tgt"].copy())
    k = 5 if name.startswith("k5") else 44
    out = metrics_ref.intersection_and_union(p, t, k, 255)
    assert np.array_equal(np.stack([o.numpy() for o in out]), z[f"c_{name}_out"])
    assert np.array_equal(p.numpy(), z[f"c_{name}_pred_after"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", METRIC)
def test_train_metric_pc_cuda_vs_reference_golden(name):
    from dropclip_b200.metrics import trainMetricPC
    z = gio.load("metrics.npz")
    preds, gts, thr, sig = _metric_inputs(name)
    p_in = [p.clone().cuda() for p in preds]
    iou, precs = trainMetricPC(p_in, [g.clone().cuda() for g in gts], threshold=thr, sigmoid=sig)
    assert iou.is_cuda and iou.dtype == torch.float32 and len(precs) == 3
    got = np.array([float(iou)] + [float(x) for x in precs])
    assert np.allclose(got, z[f"m_{name}_out"], rtol=1e-5, atol=1e-6)
    after = np.concatenate([p.reshape(-1).float().cpu().numpy() for p in p_in])
    assert np.array_equal(after, z[f"m_{name}_pred_after"], equal_nan=True)  # in-place binarisation (or none with sigmoid)
    for a, b in zip(p_in, preds):
        assert a.dtype == b.dtype and a.shape == b.shape


@pytest.mark.gpu
def test_binary_iou_counts_exact_vs_oracle_large():
    from dropclip_b200.metrics import binary_iou_counts
    rng = np.random.default_rng(3)
    sizes = [200_000, 1, 0, 65_537, 12_345]
    preds = [torch.from_numpy(rng.random(n, dtype=np.float32)) for n in sizes]
    gts = [torch.from_numpy(rng.random(n) < 0.3) for n in sizes]
    inter, union = binary_iou_counts([p.cuda() for p in preds], [g.cuda() for g in gts], 0.35, in_place=False)
    want_i = [int(((p >= 0.35) & g).sum()) for p, g in zip(preds, gts)]
    want_u = [int(((p >= 0.35) | g).sum()) for p, g in zip(preds, gts)]
    assert inter.cpu().tolist() == want_i and union.cpu().tolist() == want_u
    assert trainMetricPC_empty() == (0.0, [0.0, 0.0, 0.0])


def trainMetricPC_empty():
    from dropclip_b200.metrics import trainMetricPC
    return trainMetricPC([], [])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CLASS)
def test_intersection_and_union_cuda_vs_reference_golden(name):
    from dropclip_b200.metrics import intersectionAndUnionGPU
    z = gio.load("metrics.npz")
    p = torch.from_numpy(z[f"c_{name}_pred"].copy()).cuda()
    t = torch.from_numpy(z[f"c_{name}_tgt"].copy()).cuda()
    k = 5 if name.startswith("k5") else 44
    ai, au, at = intersectionAndUnionGPU(p, t, k, 255)
    assert ai.is_cuda and ai.dtype == torch.float32
    assert np.array_equal(np.stack([ai.cpu().numpy(), au.cpu().numpy(), at.cpu().numpy()]), z[f"c_{name}_out"])
    assert np.array_equal(p.cpu().numpy(), z[f"c_{name}_pred_after"])  # output[target == ignore] = ignore, in place
    with pytest.raises(AssertionError):
        intersectionAndUnionGPU(p.view(-1)[:10], t.view(-1)[:11], k)


def test_metrics_refuse_cpu_tensors():
    from dropclip_b200.metrics import intersectionAndUnionGPU
    with pytest.raises(RuntimeError):
        intersectionAndUnionGPU(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64), 3)
