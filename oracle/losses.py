"""Oracle for the losses on the hot path.  Test infrastructure only.

* ``cross_entropy``: what ``nn.CrossEntropyLoss(ignore_index=255)`` computes
  (scripts/train_fastscnn.py:132), restated from the definition (not by calling
  ``F.cross_entropy``) so that the two can be checked against each other.
* ``ohem``: ``losses/ohem_loss.py:10-21``.
"""
from math import log

import torch
import torch.nn.functional as F


def cross_entropy(logits, target, ignore_index=255, reduction='mean'):
    """-log_softmax(logits)[target] over non-ignored pixels.  ``mean`` divides by the
    number of non-ignored pixels (NaN when there are none); ``none`` gives 0 at
    ignored pixels."""
    logp = logits.float().log_softmax(dim=1)
    valid = target != ignore_index
    t = torch.where(valid, target, torch.zeros_like(target))
    nll = -logp.gather(1, t.unsqueeze(1)).squeeze(1)
    nll = torch.where(valid, nll, torch.zeros_like(nll))
    if reduction == 'none':
        return nll
    if reduction == 'sum':
        return nll.sum()
    return nll.sum() / valid.sum()


def cross_entropy_grad(logits, target, ignore_index=255):
    """d(mean CE)/d logits = (softmax - onehot) / n_valid on valid pixels, 0 elsewhere."""
    p = logits.float().softmax(dim=1)
    valid = target != ignore_index
    t = torch.where(valid, target, torch.zeros_like(target))
    p.scatter_add_(1, t.unsqueeze(1), -torch.ones_like(p[:, :1]))
    return p * (valid.unsqueeze(1).float() / valid.sum())


def ohem(logits, target, ignore_index=-100, thresh_loss=-log(0.7), numel_frac=0.01):
    """``ohem_loss`` losses/ohem_loss.py:10-21: per-pixel CE (0 at ignored pixels, which
    still count in ``numel``), sorted descending; if the n-th largest exceeds the
    threshold keep everything above the threshold, else the top n; mean."""
    loss = cross_entropy(logits, target, ignore_index, 'none').flatten()
    n = int(loss.numel() * numel_frac)
    loss, _ = torch.sort(loss, descending=True)
    if loss[n] > thresh_loss:
        return loss[loss > thresh_loss].mean()
    return loss[:n].mean()
