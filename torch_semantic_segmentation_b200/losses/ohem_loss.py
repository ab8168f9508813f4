"""Online hard example mining loss, drop-in for ``torch_semantic_segmentation.losses.ohem_loss``
(reference: losses/ohem_loss.py) -- the loss both reference scripts train with
(scripts/train_fastscnn.py:131,137; scripts/contextnet/train_contextnet.py:122).

Same signature and value: per-pixel cross-entropy (0 at ignored pixels, which still count in
``numel`` like in the reference), ``n = int(numel * numel_frac)``; if the (n+1)-th largest loss
exceeds ``thresh_loss`` the mean of all losses above the threshold, else the mean of the n largest.
The reference sorts all ~7 M losses and synchronises with the host for the ``if``; here the
(n+1)-th largest value comes from a 3-pass radix select and the case decision, the loss and the
per-pixel gradient weights stay on the device (``functional.OhemCrossEntropy``).  Pixels whose loss
ties with the selected order statistic share the remaining top-n slots evenly (the reference's sort
keeps an arbitrary subset of them: same loss value, equally valid subgradient).
"""
from math import log

from torch import nn

from .. import functional as Fn

__all__ = ['ohem_loss', 'OHEMLoss']


def ohem_loss(input, target, ignore_index=-100, thresh_loss=-log(0.7), numel_frac=0.01):
    if isinstance(input, Fn.DeferredLogits) and input.scores.shape[1] not in (11, 12, 19, 21):
        input = input.materialize()
    if input.dim() != 4 or target.dim() != 3:
        raise ValueError('ohem_loss expects (N,C,H,W) logits and (N,H,W) targets')
    n = input.shape[0] * input.shape[2] * input.shape[3]
    if not 0 <= int(n * numel_frac) < n:
        raise IndexError('ohem_loss: numel_frac=%r selects index %d of %d losses' % (numel_frac, int(n * numel_frac), n))
    scores = Fn.fused_head_source(input)
    if scores is not None and scores.shape[1] in (11, 12, 19, 21):
        return Fn.OhemCrossEntropy.apply(scores, target, ignore_index, thresh_loss, numel_frac,
                                         (input.shape[2], input.shape[3]))
    return Fn.OhemCrossEntropy.apply(input, target, ignore_index, thresh_loss, numel_frac, None)


class OHEMLoss(nn.Module):
    accepts_deferred_logits = True       # functional.DeferredLogits

    def __init__(self, ignore_index=-100, thresh_loss=-log(0.7), numel_frac=0.01):
        super().__init__()
        self.ignore_index = ignore_index
        self.thresh_loss = thresh_loss
        self.numel_frac = numel_frac

    def forward(self, input, target):
        return ohem_loss(input, target, ignore_index=self.ignore_index, thresh_loss=self.thresh_loss,
                         numel_frac=self.numel_frac)
