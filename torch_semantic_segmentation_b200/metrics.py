"""Evaluation metrics on the sm_100a confusion-matrix kernel.

Stand-ins for the pytorch-ignite classes the reference wires up in engine.py:65-72
(``ConfusionMatrix`` -> ``IoU`` / ``mIoU`` / ``cmAccuracy`` / ``DiceCoefficient``), with
the ignite >= 0.4 formulas: the matrix is an int64 (C, C) histogram
of (target, argmax prediction) over pixels with 0 <= target < C, accumulated ON DEVICE;
the derived metrics are float64 with the +1e-15 guards.  In a process group the matrix is
summed with ONE int64 all-reduce.  The four reference metrics share one matrix here (the
reference builds four identical ones, engine.py:67-71).
"""
import torch
import torch.distributed as dist

from . import ops

__all__ = ['ConfusionMatrix', 'iou', 'miou', 'cm_accuracy', 'dice_coefficient', 'metrics_from_cm']


class ConfusionMatrix:
    """``update((y_pred, y))`` with y_pred (N, C, H, W) logits (argmax fused into the kernel)
    or an (N, H, W) / flat int64 prediction map; ``compute()`` -> int64 (C, C) CPU tensor."""

    def __init__(self, num_classes, device=None):
        self.num_classes = num_classes
        self.device = device
        self.cm = None
        self.num_examples = 0

    def reset(self):
        if self.cm is not None:
            self.cm.zero_()
        self.num_examples = 0

    def _ensure(self, device):
        if self.cm is None or self.cm.device != device:
            self.cm = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=device)

    def update(self, output):
        y_pred, y = output
        self._ensure(y_pred.device)
        if y_pred.dim() == y.dim() + 1:
            if y_pred.shape[1] != self.num_classes:
                raise ValueError('y_pred has %d classes, expected %d' % (y_pred.shape[1], self.num_classes))
            ops.confusion_from_logits(y_pred, y, self.cm)
        elif y_pred.shape == y.shape:
            ops.confusion_from_labels(y_pred, y, self.num_classes, self.cm)
        else:
            raise ValueError('y_pred %s and y %s do not match' % (tuple(y_pred.shape), tuple(y.shape)))
        self.num_examples += y.shape[0]

    def compute(self, sync=True):
        if self.cm is None:
            raise RuntimeError('ConfusionMatrix must have at least one example before it can be computed')
        cm = self.cm
        if sync and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            cm = cm.clone()
            dist.all_reduce(cm, op=dist.ReduceOp.SUM)      # one int64 all-reduce of C*C counts
        return cm.cpu()


def iou(cm):
    cm = cm.to(torch.float64)
    diag = cm.diag()
    return diag / (cm.sum(dim=1) + cm.sum(dim=0) - diag + 1e-15)


def miou(cm):
    return iou(cm).mean()


def cm_accuracy(cm):
    cm = cm.to(torch.float64)
    return cm.diag().sum() / (cm.sum() + 1e-15)


def dice_coefficient(cm):
    cm = cm.to(torch.float64)
    return 2.0 * cm.diag() / (cm.sum(dim=1) + cm.sum(dim=0) + 1e-15)


def metrics_from_cm(cm):
    return {'iou': iou(cm), 'miou': miou(cm), 'accuracy': cm_accuracy(cm), 'dice': dice_coefficient(cm)}
