"""ORACLE (test infrastructure): CPU restatement of the reference's metric reductions.

`train_metric_pc` follows utils/misc.py:21-50 line by line (torch on the CPU, including the in-place
binarisation of the prediction tensors); `intersection_and_union` follows utils/misc.py:186-199 without
the final `.cuda()` (the reference itself computes the histograms on the CPU with torch.histc).
Pinned against outputs of the unmodified reference in tests/golden/metrics.npz (tests/make_golden.py)."""
import torch


@torch.no_grad()
def train_metric_pc(output, target, threshold=0.35, pr_ious=(0.25, 0.5, 0.75), sigmoid=False):
    assert len(output) == len(target)
    mean_iou = 0.0
    mean_prec = [0.0] * len(pr_ious)
    count = 1e-6
    for pred, gt in zip(output, target):
        count += 1
        pred = torch.sigmoid(pred).squeeze() if sigmoid else pred.squeeze()
        pred[pred < threshold] = 0.
        pred[pred >= threshold] = 1.
        inter = (pred.bool() & gt.bool()).sum()
        union = (pred.bool() | gt.bool()).sum()
        iou = inter / (union + 1e-6)
        mean_iou += iou
        for j, pr in enumerate(pr_ious):
            mean_prec[j] += (iou > pr).float()
    mean_iou /= count + 1e-6
    mean_prec = [p / count for p in mean_prec]
    return 100. * mean_iou, [100. * x for x in mean_prec]


def intersection_and_union(output, target, K, ignore_index=255):
    assert output.dim() in [1, 2, 3, 4]
    assert output.shape == target.shape
    output = output.view(-1)
    target = target.view(-1)
    output[target == ignore_index] = ignore_index
    intersection = output[output == target]
    area_intersection = torch.histc(intersection.float().cpu(), bins=K, min=0, max=K - 1)
    area_output = torch.histc(output.float().cpu(), bins=K, min=0, max=K - 1)
    area_target = torch.histc(target.float().cpu(), bins=K, min=0, max=K - 1)
    area_union = area_output + area_target - area_intersection
    return area_intersection, area_union, area_target
