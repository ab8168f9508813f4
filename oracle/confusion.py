"""Oracle for the evaluation metrics (numpy, integer-exact).  Test infrastructure only.

The reference builds these from pytorch-ignite (``engine.py:6-11,65-72``):
``ConfusionMatrix(num_classes)`` feeding ``IoU``, ``mIoU``, ``cmAccuracy`` and
``DiceCoefficient``.  ignite is NOT under /root/reference and is not pinned by the
reference (no requirements / install_requires), so this file restates the published
ignite>=0.4 algorithm:

    p  = argmax(y_pred, dim=1)              (ties -> lowest class index)
    m  = (y >= 0) & (y < C)
    cm = bincount(C * y[m] + p[m], minlength=C*C).reshape(C, C)   rows = target
    IoU  = diag / (rowsum + colsum - diag + 1e-15)                 float64
    mIoU = IoU.mean()
    acc  = diag.sum() / (cm.sum() + 1e-15)
    dice = 2 diag / (rowsum + colsum + 1e-15)

PARITY UNPINNED by the reference itself for this part (no golden vectors, no source);
``tests/test_oracle.py`` cross-checks it against scikit-learn.
"""
import numpy as np


def argmax_classes(logits):
    """``torch.argmax(y_pred, dim=1)`` semantics on a numpy (N, C, H, W) array: first
    maximal index; a NaN counts as the maximum (torch behaviour)."""
    x = np.asarray(logits, dtype=np.float32)
    x = np.where(np.isnan(x), np.inf, x)
    return np.argmax(x, axis=1).astype(np.int64)


def confusion_matrix(pred, target, num_classes=19):
    """19x19 int64 histogram of (target, prediction) over valid targets."""
    pred = np.asarray(pred).astype(np.int64).ravel()
    target = np.asarray(target).astype(np.int64).ravel()
    m = (target >= 0) & (target < num_classes)
    idx = num_classes * target[m] + pred[m]
    return np.bincount(idx, minlength=num_classes * num_classes) \
        .reshape(num_classes, num_classes).astype(np.int64)


def metrics(cm):
    """IoU / mIoU / accuracy / dice from a confusion matrix, float64, ignite formulas."""
    cm = np.asarray(cm, dtype=np.float64)
    diag = np.diag(cm)
    rows, cols = cm.sum(axis=1), cm.sum(axis=0)
    iou = diag / (rows + cols - diag + 1e-15)
    return {
        'iou': iou,
        'miou': iou.mean(),
        'accuracy': diag.sum() / (cm.sum() + 1e-15),
        'dice': 2.0 * diag / (rows + cols + 1e-15),
    }
