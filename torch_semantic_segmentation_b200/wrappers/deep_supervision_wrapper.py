"""``DeepSupervisionWrapper``, drop-in for ``torch_semantic_segmentation.wrappers``
(reference: wrappers/deep_supervision_wrapper.py:12-43).

In training mode the auxiliary heads run on the outputs of the given sub-modules of the wrapped
model and ``forward`` returns ``(output, [aux_0, aux_1, ...])``; in eval mode only the model's
output.  Same constructor, same ``state_dict`` layout (``module.*``, ``auxiliary.<i>.*``), so the
reference's checkpoints of wrapped models load with ``strict=True``
(scripts/train_fastscnn.py:108-121).  The reference installs and removes forward hooks on every call; here one
persistent hook per tapped sub-module is installed at construction and armed for the duration of a training
forward.  The tapped sub-modules of this package's models (``model.downsample`` / ``model.features``) emit
NCHW-logical tensors exactly like the reference's.
``AuxiliaryHead`` is the head the reference script builds inline
(``nn.Sequential(Classifier(C, classes), nn.Upsample(scale_factor=s, bilinear, align_corners=True))``)
on the sm_100a kernels, with the up-sampling fused into the loss when the loss is this package's.
"""
from typing import List, Tuple

import torch
from torch import nn

from .. import functional as Fn
from .. import ops

__all__ = ['DeepSupervisionWrapper', 'AuxiliaryHead']


class _Tap:
    """Persistent forward hook on one tapped sub-module.  While the wrapper's training forward is running
    (``armed``) it runs the auxiliary head on the sub-module's output right where the reference's per-call hook
    does -- inside the wrapped model's forward, so the heads' dropout draws come in the same order -- and keeps the
    result until the wrapper collects it.  Outside of that window (eval mode, or the wrapped model called on its
    own) it does nothing."""

    __slots__ = ('head', 'armed', 'result')

    def __init__(self, head):
        self.head, self.armed, self.result = head, False, None

    def __call__(self, _layer, _args, feature):
        if self.armed:
            self.result = self.head(feature)

    def collect(self):
        result, self.result, self.armed = self.result, None, False
        return result


class DeepSupervisionWrapper(nn.Module):
    """``DeepSupervisionWrapper(model, [(tapped_submodule, head), ...])``.  The hooks are installed once, here
    (a captured CUDA graph and the eager loop see the same module tree; nothing is registered per step)."""

    def __init__(self, module: nn.Module, auxiliary_modules: List[Tuple[nn.Module, nn.Module]]):
        super().__init__()
        self.module = module
        pairs = list(auxiliary_modules)
        self.auxiliary = nn.ModuleList(head for _, head in pairs)
        # not registered as children: the tapped layers already live inside ``module`` (state_dict = module.* + auxiliary.*)
        self.layers = [layer for layer, _ in pairs]
        self._taps = [_Tap(head) for head in self.auxiliary]
        for layer, tap in zip(self.layers, self._taps):
            layer.register_forward_hook(tap)

    def set_compute_dtype(self, dtype, pw_impl=None):
        from ..nn.blocks import set_compute_dtype
        if pw_impl is None:
            pw_impl = 1 if dtype == torch.bfloat16 else 0
        set_compute_dtype(self, dtype, pw_impl)
        return self

    def forward(self, input):
        if not self.training:
            return self.module(input)
        for tap in self._taps:
            tap.armed, tap.result = True, None
        try:
            output = self.module(input)
        finally:
            extra = [tap.collect() for tap in self._taps]
        return output, extra


class AuxiliaryHead(nn.Sequential):
    """``nn.Sequential(classifier, nn.Upsample(scale_factor, 'bilinear', align_corners=True))`` with the
    reference's child indices (``0.*`` = the classifier's parameters, ``1`` = the parameter-free
    up-sampling).  The up-sampling runs on ``tss_upsample_logits_fwd`` and remembers its source so
    that the loss can take the fused head."""

    def __init__(self, classifier, scale_factor):
        super().__init__(classifier, nn.Upsample(scale_factor=scale_factor, mode='bilinear', align_corners=True))
        self.scale_factor = int(scale_factor)

    def forward(self, input):
        scores = ops.as_nhwc(self[0](input))
        s = self.scale_factor
        logits = Fn.UpsampleLogits.apply(scores, scores.shape[2] * s, scores.shape[3] * s)
        return Fn.attach_head(logits, scores)
