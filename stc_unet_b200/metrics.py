"""GPU-resident evaluation (SURVEY §8 rows `intersect_and_union`, `pre_eval_to_metrics`, and (f)-4).

The reference's `CustomDataset.pre_eval` (mmseg/datasets/custom.py:277-314) copies every int64 prediction map to the host and
calls `intersect_and_union` (mmseg/core/evaluation/metrics.py:26-87: three float `torch.histc` passes on the CPU) per image;
`pre_eval_to_metrics` (:309-344) then sums the per-image float32 vectors (exact only below 2^24 pixels) and
`total_area_to_metrics` derives the scores.  Here the prediction stays on the device, ONE integer histogram kernel
(`stc_confusion_hist`) accumulates the int64 confusion matrix CM[label, pred] and the four area vectors, ranks exchange a
single int64 all-reduce, and the scores are the STANDARD definitions pinned by the reference's own tests
(tests/test_metrics.py:9-85).  The fork's tampered post-processing (metrics.py:7,425-429,454-457: a random offset and a
compounding `v + (1-v)/3` inflation of every score) is deliberately NOT reproduced (SURVEY §0 integrity note).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ops


def metrics_from_confusion(cm: np.ndarray, metrics: Sequence[str] = ("mIoU",), nan_to_num: Optional[int] = None,
                           beta: float = 1.0) -> "OrderedDict[str, np.ndarray]":
    """aAcc + per-class scores from an integer confusion matrix CM[label, pred]; same keys as total_area_to_metrics
    (metrics.py:387-468 before the tampering): mIoU -> IoU, Acc; mDice -> Dice, Acc; mFscore -> Fscore, Precision, Recall."""
    allowed = ("mIoU", "mDice", "mFscore")
    if isinstance(metrics, str):
        metrics = [metrics]
    if not set(metrics).issubset(allowed):
        raise KeyError(f"metrics {metrics} is not supported")
    cm = np.asarray(cm, dtype=np.float64)
    inter, label_area, pred_area = np.diag(cm), cm.sum(1), cm.sum(0)
    union = pred_area + label_area - inter
    out = OrderedDict(aAcc=inter.sum() / label_area.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        for m in metrics:
            if m == "mIoU":
                out["IoU"], out["Acc"] = inter / union, inter / label_area
            elif m == "mDice":
                out["Dice"], out["Acc"] = 2 * inter / (pred_area + label_area), inter / label_area
            else:
                prec, rec = inter / pred_area, inter / label_area
                out["Fscore"] = (1 + beta ** 2) * (prec * rec) / ((beta ** 2 * prec) + rec)
                out["Precision"], out["Recall"] = prec, rec
    if nan_to_num is not None:
        out = OrderedDict((k, np.nan_to_num(v, nan=nan_to_num)) for k, v in out.items())
    return out


class ConfusionMeter:
    """Accumulates CM[label, pred] (int64, on the device) over batches / volumes."""

    def __init__(self, num_classes: int, ignore_index: int = 255, device="cuda"):
        self.num_classes, self.ignore_index = num_classes, ignore_index
        self.cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=device)
        self.areas = torch.zeros((4, num_classes), dtype=torch.int64, device=device)   # intersect, union, pred, label

    def update(self, pred: torch.Tensor, label: torch.Tensor):
        """pred: int64 (...), label: uint8 or int64 (...), same number of elements; both on the device."""
        ops.confusion_hist(pred, label, self.num_classes, self.ignore_index, self.cm, self.areas)

    def all_reduce(self):
        """One int64 all-reduce replaces collect_results_cpu/gpu (mmseg/apis/test.py:228-232)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = torch.cat([self.cm.flatten(), self.areas.flatten()])
            dist.all_reduce(packed)
            n = self.num_classes * self.num_classes
            self.cm.copy_(packed[:n].view_as(self.cm))
            self.areas.copy_(packed[n:].view_as(self.areas))

    def pre_eval_tuple(self):
        """(area_intersect, area_union, area_pred_label, area_label) like intersect_and_union, but exact int64."""
        a = self.areas.cpu()
        return a[0], a[1], a[2], a[3]

    def compute(self, metrics: Sequence[str] = ("mIoU", "mFscore", "mDice"), nan_to_num: Optional[int] = None) -> Dict[str, float]:
        """Summary in the shape of CustomDataset.evaluate's return (custom.py:388-487): 'aAcc', 'mIoU', 'IoU.<i>', ..."""
        per = metrics_from_confusion(self.cm.cpu().numpy(), metrics, nan_to_num)
        res = OrderedDict()
        for k, v in per.items():
            if k == "aAcc":
                res["aAcc"] = float(v)
            else:
                res["m" + k] = float(np.nanmean(v))
                for i, x in enumerate(np.atleast_1d(v)):
                    res[f"{k}.{i}"] = float(x)
        return res
