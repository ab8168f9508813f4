"""ContextNet on hand-written sm_100a kernels, drop-in for
``torch_semantic_segmentation.models.contextnet`` (reference: models/contextnet.py).

Same factories (``contextnet12/14/18`` = input shrink 2/4/8, contextnet.py:13-25), same
constructor, same child names (``spatial``, ``context``, ``feature_fusion``, ``classifier``), same
``state_dict`` keys/shapes and the same random init under the same seed.  Input: NCHW float32
``(N, 3, H, W)`` (any size, like the reference: the fusion module resizes the context map to the
spatial map, contextnet.py:119-121); output NCHW-contiguous ``(N, out_channels, 8*(H/8), 8*(W/8))``.
"""
import torch
from torch import nn

from .. import ops
from .. import functional as Fn
from ..nn.blocks import (ConvBNBlock, BottleneckBlock, ClassScores, Dropout, hands_over_to_pointwise, observed,
                         set_compute_dtype)

__all__ = ['ContextNet', 'contextnet12', 'contextnet14', 'contextnet18']


def contextnet12(in_channels, out_channels):
    return ContextNet(in_channels, out_channels, scale_factor=2)


def contextnet14(in_channels, out_channels):
    return ContextNet(in_channels, out_channels, scale_factor=4)


def contextnet18(in_channels, out_channels):
    return ContextNet(in_channels, out_channels, scale_factor=8)


def ConvBlock(in_channels, out_channels, kernel_size, padding=0, stride=1, use_relu=True):
    """reference: models/contextnet.py:168-177."""
    return ConvBNBlock(in_channels, out_channels, kernel_size, stride, padding, 1, 1, use_relu)


class _Chain(nn.Sequential):
    """``nn.Sequential`` (same child indices, same state_dict keys) of conv+BN blocks in which every block is the only
    reader of its predecessor's output: with functional.FUSE_BNRED_EXT the predecessor's BatchNorm-backward reduction
    is folded into each block's dgrad (the spatial branch holds the largest activations of the network)."""

    @staticmethod
    def _hands_over(dw, pw):
        return hands_over_to_pointwise(dw, pw)

    @staticmethod
    def _hands_over_to_dw(blk, dw):
        """``blk`` (any conv+BN block) may leave its BatchNorm + ReLU to the depthwise block ``dw`` that follows:
        functional.FUSE_BNIN."""
        if not (Fn.FUSE_BNIN and isinstance(blk, ConvBNBlock) and isinstance(dw, ConvBNBlock) and blk.training
                and torch.is_grad_enabled() and not observed(blk)):
            return False
        c = dw[0]
        return bool(c.groups == c.in_channels and c.groups > 1 and c.kernel_size == (3, 3) and c.dilation[0] == 1
                    and c.stride[0] in (1, 2) and (c.in_channels % 32 == 0 or c.in_channels % 48 == 0)
                    and blk[1].track_running_stats and getattr(blk[1], '_tss_sync', None) is None
                    and getattr(dw[1], '_tss_sync', None) is None)

    def forward(self, input):
        x, prev = input, None
        modules = list(self)
        for i, module in enumerate(modules):
            if isinstance(module, ConvBNBlock):
                nxt = modules[i + 1] if i + 1 < len(modules) else None
                x = module(x, sole_consumer=Fn.FUSE_BNRED_EXT and isinstance(prev, ConvBNBlock),
                           defer_apply=self._hands_over(module, nxt) or self._hands_over_to_dw(module, nxt))
            else:
                x = module(x)
            prev = module
        return x


def DWConvBlock(in_channels, out_channels, kernel_size, padding=0, stride=1, dilation=1, use_relu=True):
    """reference: models/contextnet.py:150-165."""
    if in_channels != out_channels:
        raise ValueError("input and output channels must be the same in depthwise convolution")
    return ConvBNBlock(in_channels, out_channels, kernel_size, stride, padding, dilation, in_channels, use_relu)


def LinearBottleneck(in_channels, out_channels, num_blocks, expansion=6, stride=1):
    """reference: models/contextnet.py:90-101."""
    layers = [BottleneckBlock(in_channels, out_channels, stride=stride, expansion=expansion)]
    for _ in range(1, num_blocks):
        layers.append(BottleneckBlock(out_channels, out_channels, expansion=expansion))
    return nn.Sequential(*layers)


def Classifier(in_channels, out_channels):
    """reference: models/contextnet.py:79-87."""
    return _Chain(
        DWConvBlock(in_channels, in_channels, 3, padding=1),
        ConvBlock(in_channels, in_channels, 1),
        DWConvBlock(in_channels, in_channels, 3, padding=1),
        ConvBlock(in_channels, in_channels, 1),
        Dropout(p=0.1),
        ClassScores(in_channels, out_channels),
    )


class FeatureFusionModule(nn.Module):
    """reference: models/contextnet.py:104-126."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        lowres_channels, highres_channels = in_channels
        self.lowres = nn.Sequential(
            DWConvBlock(lowres_channels, lowres_channels, kernel_size=3, padding=4, dilation=4),
            ConvBlock(lowres_channels, out_channels, 1, use_relu=False),
        )
        self.highres = ConvBlock(highres_channels, out_channels, 1, use_relu=False)

    def forward(self, lowres, highres):
        lowres = ops.as_nhwc(lowres)
        x = Fn.Bilinear.apply(lowres, highres.shape[2], highres.shape[3])
        x = self.lowres[0](x, defer_apply=hands_over_to_pointwise(self.lowres[0], self.lowres[1]))
        high = self.highres(highres)
        # relu(lowres + highres): add and ReLU fused into the low-res branch's BatchNorm apply
        return self.lowres[1](x, residual=high, relu=True)


class ContextNet(nn.Module):
    """reference: models/contextnet.py:28-76."""

    scale_factor: int = 4

    def __init__(self, in_channels, out_channels, scale_factor=4):
        super().__init__()
        self.scale_factor = scale_factor
        self.spatial = _Chain(
            ConvBlock(in_channels, 32, 3, padding=1, stride=2),
            DWConvBlock(32, 32, kernel_size=3, padding=1, stride=2),
            ConvBlock(32, 64, 1),
            DWConvBlock(64, 64, kernel_size=3, padding=1, stride=2),
            ConvBlock(64, 128, 1),
            DWConvBlock(128, 128, kernel_size=3, padding=1, stride=1),
            ConvBlock(128, 128, 1),
        )
        self.context = nn.Sequential(
            ConvBlock(in_channels, 32, 3, padding=1, stride=2),
            BottleneckBlock(32, 32, expansion=1),
            BottleneckBlock(32, 32, expansion=6),
            LinearBottleneck(32, 48, 3, stride=2),
            LinearBottleneck(48, 64, 3, stride=2),
            LinearBottleneck(64, 96, 2),
            LinearBottleneck(96, 128, 2),
            ConvBlock(128, 128, 3, padding=1),
        )
        self.feature_fusion = FeatureFusionModule((128, 128), 128)
        self.classifier = Classifier(128, out_channels)
        self.defer_logits = False        # functional.DeferredLogits in training mode (set by the trainer)

    def set_compute_dtype(self, dtype, pw_impl=None):
        """float32 (verification mode, default) or bfloat16 (tcgen05 pointwise convolutions)."""
        if pw_impl is None:
            pw_impl = 1 if dtype == torch.bfloat16 else 0
        set_compute_dtype(self, dtype, pw_impl)
        return self

    def forward(self, input):
        if input.dim() != 4:
            raise RuntimeError('ContextNet expects (N, C, H, W) input, got %s' % (tuple(input.shape),))
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        spatial = self.spatial(input)
        # F.interpolate(scale_factor=1/s) sizes the output as floor(in / s) (contextnet.py:65-67)
        H, W = input.shape[2], input.shape[3]
        context = ops.bilinear_nchw_f32(input, int(H * (1.0 / self.scale_factor)), int(W * (1.0 / self.scale_factor)))
        context = self.context(context)
        fusion = self.feature_fusion(context, spatial)
        classes = self.classifier(fusion)
        classes = ops.as_nhwc(classes)
        return Fn.model_output(self, classes, classes.shape[2] * 8, classes.shape[3] * 8)
