"""``DeepSupervisionWrapper``, drop-in for ``torch_semantic_segmentation.wrappers``
(reference: wrappers/deep_supervision_wrapper.py:12-43).

In training mode the auxiliary heads run on the outputs of the given sub-modules of the wrapped
model and ``forward`` returns ``(output, [aux_0, aux_1, ...])``; in eval mode only the model's
output.  Same constructor, same ``state_dict`` layout (``module.*``, ``auxiliary.<i>.*``), so the
reference's checkpoints of wrapped models load with ``strict=True``
(scripts/train_fastscnn.py:108-121).  The reference installs forward hooks for every call; the
hooked sub-modules of this package's models (``model.downsample`` / ``model.features``) emit
NCHW-logical tensors exactly like the reference's, so the same mechanism is used here.
``AuxiliaryHead`` is the head the reference script builds inline
(``nn.Sequential(Classifier(C, classes), nn.Upsample(scale_factor=s, bilinear, align_corners=True))``)
on the sm_100a kernels, with the up-sampling fused into the loss when the loss is this package's.
"""
from functools import partial
from typing import List, Tuple

from torch import nn

from .. import functional as Fn
from .. import ops

__all__ = ['DeepSupervisionWrapper', 'AuxiliaryHead']


class DeepSupervisionWrapper(nn.Module):

    def __init__(self, module: nn.Module, auxiliary_modules: List[Tuple[nn.Module, nn.Module]]):
        super().__init__()
        self.module = module
        self.layers = [layer for layer, _module in auxiliary_modules]
        self.auxiliary = nn.ModuleList([module for _layer, module in auxiliary_modules])

    def set_compute_dtype(self, dtype, pw_impl=None):
        import torch
        from ..nn.blocks import set_compute_dtype
        if pw_impl is None:
            pw_impl = 1 if dtype == torch.bfloat16 else 0
        set_compute_dtype(self, dtype, pw_impl)
        return self

    def forward(self, input):
        if self.training:
            aux_outputs = [None for _ in range(len(self.layers))]
            hooks = []
            for id, (layer, auxiliary) in enumerate(zip(self.layers, self.auxiliary)):
                hook_fn = partial(auxiliary_hook, aux_outputs=aux_outputs, aux_id=id, auxiliary_module=auxiliary)
                hooks.append(layer.register_forward_hook(hook_fn))
            try:
                output = self.module(input)
            finally:
                for hook in hooks:
                    hook.remove()
            return output, aux_outputs
        return self.module(input)


def auxiliary_hook(_module, _input, output, aux_outputs, aux_id, auxiliary_module):
    aux_outputs[aux_id] = auxiliary_module(output)


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
