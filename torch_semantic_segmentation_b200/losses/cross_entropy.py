"""Softmax cross-entropy with ``ignore_index`` on the fused sm_100a kernel.

Drop-in for the ``nn.CrossEntropyLoss(ignore_index=255)`` the reference trains with
(scripts/train_fastscnn.py:132): logits ``(N, C, H, W)`` (fp32 or bf16, NCHW), target
``(N, H, W)`` int64; ``reduction='mean'`` divides by the number of non-ignored pixels
(NaN if there are none) and ``reduction='none'`` returns 0 at ignored pixels.
Forward and gradient are computed in ONE pass over the logits.  When ``input`` is the untouched
output of one of this package's models, the loss is computed from the 1/8-resolution class scores
it was interpolated from (``functional.UpsampleCrossEntropy``): same value up to fp32 rounding,
without reading the full-resolution logits or materialising their gradient.
"""
import torch
from torch import nn

from .. import ops
from .. import functional as Fn

__all__ = ['CrossEntropyLoss', 'cross_entropy']


def cross_entropy(input, target, ignore_index=-100, reduction='mean', fused_head=True):
    if isinstance(input, Fn.DeferredLogits) and not (reduction == 'mean' and fused_head
                                                     and input.scores.shape[1] in (11, 12, 19, 21)):
        input = input.materialize()          # this call needs the real tensor
    if input.dim() != 4 or target.dim() != 3:
        raise ValueError('cross_entropy expects (N,C,H,W) logits and (N,H,W) targets')
    if reduction == 'mean':
        scores = Fn.fused_head_source(input) if fused_head else None
        if scores is not None and scores.shape[1] in (11, 12, 19, 21):
            return Fn.UpsampleCrossEntropy.apply(scores, target, ignore_index, input.shape[2], input.shape[3])
        return Fn.CrossEntropy.apply(input, target, ignore_index)
    if reduction in ('none', 'sum'):
        if torch.is_grad_enabled() and input.requires_grad:
            pixel = Fn.PixelCrossEntropy.apply(input, target, ignore_index)
        else:
            pixel = ops.ce_forward(input, target, ignore_index, want_grad=False, want_pixel_loss=True)[2]
        return pixel if reduction == 'none' else pixel.sum()
    raise ValueError('unknown reduction %r' % reduction)


class CrossEntropyLoss(nn.Module):
    accepts_deferred_logits = True       # functional.DeferredLogits

    def __init__(self, ignore_index=-100, reduction='mean', fused_head=True):
        super().__init__()
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.fused_head = fused_head

    def forward(self, input, target):
        return cross_entropy(input, target, self.ignore_index, self.reduction, self.fused_head)
