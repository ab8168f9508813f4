"""Functional restatement of the reference's file-local conv blocks.

Oracle / test infrastructure only (see ``oracle/__init__.py``).

Every function takes ``sd`` (a dict name -> fp32 CPU tensor with the reference's
``state_dict`` keys; parameters may ``require_grad``) and ``p`` (the key prefix of
the ``nn.Sequential`` the reference builds) and applies the same stock torch ops.
BatchNorm buffers in ``sd`` are updated in place in training mode exactly as
``nn.BatchNorm2d`` does (momentum 0.1, eps 1e-5, unbiased running_var).
"""
import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.1
BN_EPS = 1e-5


def batch_norm(sd, p, x, training):
    """``nn.BatchNorm2d(C)`` with default arguments (fastscnn.py:169,181,193,195)."""
    if training:
        sd[p + '.num_batches_tracked'] += 1
    return F.batch_norm(
        x, sd[p + '.running_mean'], sd[p + '.running_var'],
        sd[p + '.weight'], sd[p + '.bias'],
        training, BN_MOMENTUM, BN_EPS)


def conv_block(sd, p, x, training, stride=1, padding=0, dilation=1, relu=True):
    """``Conv2dBlock`` fastscnn.py:164-173 / ``ConvBlock`` contextnet.py:168-177:
    Conv2d(bias=False) -> BatchNorm2d -> optional ReLU."""
    x = F.conv2d(x, sd[p + '.0.weight'], None, stride, padding, dilation, 1)
    x = batch_norm(sd, p + '.1', x, training)
    return F.relu(x) if relu else x


def dw_block(sd, p, x, training, stride=1, padding=0, dilation=1, relu=True):
    """``DWConv2dBlock`` fastscnn.py:176-185 / ``DWConvBlock`` contextnet.py:150-165:
    depthwise Conv2d(groups=C, bias=False) -> BatchNorm2d -> optional ReLU."""
    w = sd[p + '.0.weight']
    x = F.conv2d(x, w, None, stride, padding, dilation, w.shape[0])
    x = batch_norm(sd, p + '.1', x, training)
    return F.relu(x) if relu else x


def ds_block(sd, p, x, training, stride=1, padding=0, dilation=1, relu=True):
    """``DSConv2dBlock`` fastscnn.py:188-199: dw3x3 -> BN (no ReLU) -> 1x1 -> BN -> ReLU."""
    w = sd[p + '.0.weight']
    x = F.conv2d(x, w, None, stride, padding, dilation, w.shape[0])
    x = batch_norm(sd, p + '.1', x, training)
    x = F.conv2d(x, sd[p + '.2.weight'])
    x = batch_norm(sd, p + '.3', x, training)
    return F.relu(x) if relu else x


def bottleneck(sd, p, x, training, stride=1):
    """``BottleneckBlock`` fastscnn.py:138-161 / contextnet.py:129-147: pw-expand(+ReLU)
    -> dw3x3(stride, +ReLU) -> pw-project (linear) -> +input iff shapes equal -> ReLU
    (the trailing ReLU is unconditional in the reference)."""
    y = conv_block(sd, p + '.conv1', x, training)
    y = dw_block(sd, p + '.conv2', y, training, stride=stride, padding=1)
    y = conv_block(sd, p + '.conv3', y, training, relu=False)
    if y.shape == x.shape:
        y = y + x
    return F.relu(y)


def upsample(x, size=None, scale_factor=None):
    """Bilinear ``align_corners=True`` (fastscnn.py:63-64,74,119-120; contextnet.py:65-76,119-121)."""
    return F.interpolate(x, size=size, scale_factor=scale_factor,
                         mode='bilinear', align_corners=True)
