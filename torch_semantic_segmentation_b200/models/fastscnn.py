"""Fast-SCNN on hand-written sm_100a kernels, drop-in for
``torch_semantic_segmentation.models.fastscnn`` (reference: models/fastscnn.py).

Same constructor signature, same child names (``downsample``, ``features``, ``fusion``,
``classifier``), same ``state_dict`` keys/shapes and the same random init under the same
seed, so reference checkpoints load with ``strict=True`` and forward hooks on
``model.downsample`` / ``model.features`` (``DeepSupervisionWrapper``) keep working.
Input: NCHW float32 ``(N, 3, H, W)``, H and W multiples of 32; output: NCHW-contiguous
``(N, out_channels, H, W)`` in the compute dtype.
"""
import torch
from torch import nn

from .. import ops
from .. import functional as Fn
from ..nn.blocks import (Conv2dBlock, DWConv2dBlock, DSConv2dBlock, DSConvBNBlock, BottleneckBlock, ClassScores, Dropout,
                         hands_over_to_pointwise, observed, set_compute_dtype)

__all__ = ['FastSCNN', 'fastscnn', 'Classifier']


def fastscnn(in_channels, out_channels):
    return FastSCNN(in_channels, out_channels)


class FastSCNN(nn.Module):
    """reference: models/fastscnn.py:15-64."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.downsample = _DownsampleChain(
            Conv2dBlock(in_channels, 32, kernel_size=3, padding=1, stride=2),
            DSConv2dBlock(32, 48, kernel_size=3, padding=1, stride=2),
            DSConv2dBlock(48, 64, kernel_size=3, padding=1, stride=2),
        )
        # inside `downsample` each block is the only reader of its predecessor's output (forward hooks are
        # documented on `downsample` / `features` as a whole, wrappers/deep_supervision_wrapper.py:28-37)
        self.downsample[1].input_sole_consumer = True
        self.downsample[2].input_sole_consumer = True
        self.features = nn.Sequential(
            BottleneckModule(64, 64, expansion=6, repeats=3, stride=2),
            BottleneckModule(64, 96, expansion=6, repeats=3, stride=2),
            BottleneckModule(96, 128, expansion=6, repeats=3, stride=1),
            PyramidPoolingModule(128, 128),
        )
        self.fusion = FeatureFusionModule((128, 64), 128, scale_factor=4)
        self.classifier = Classifier(128, out_channels)
        self.defer_logits = False        # functional.DeferredLogits in training mode (set by the trainer, see there)

    def set_compute_dtype(self, dtype, pw_impl=None):
        """float32 (verification mode, default) or bfloat16 (the B200 production path: pointwise
        convolutions on tcgen05 tensor cores unless ``pw_impl=0`` asks for the SIMT kernels)."""
        if pw_impl is None:
            pw_impl = 1 if dtype == torch.bfloat16 else 0
        set_compute_dtype(self, dtype, pw_impl)
        return self

    def forward(self, input):
        if input.dim() != 4 or input.shape[2] % 32 or input.shape[3] % 32:
            raise RuntimeError('FastSCNN expects (N, C, H, W) input with H and W multiples of 32, got %s'
                               % (tuple(input.shape),))
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        downsample = self.downsample(input)
        features = self.features(downsample)
        fusion = self.fusion(features, downsample)
        classes = self.classifier(fusion)
        classes = ops.as_nhwc(classes)
        return Fn.model_output(self, classes, classes.shape[2] * 8, classes.shape[3] * 8)


class _DownsampleChain(nn.Sequential):
    """``nn.Sequential`` (same child indices and state_dict keys) for learning-to-downsample, in which every block is the
    only reader of its predecessor's output: with functional.FUSE_BNIN a block leaves its BatchNorm + ReLU to the
    depthwise conv that follows (the stem's output -- 32 channels at 1/2 resolution, the largest activation of the
    network -- is then never materialised).  Forward hooks on the chain as a whole see the last block's real output."""

    def forward(self, input):
        x = input
        modules = list(self)
        for i, module in enumerate(modules):
            nxt = modules[i + 1] if i + 1 < len(modules) else None
            last_bn = module[3] if isinstance(module, DSConvBNBlock) else module[1]      # the BatchNorm that would be deferred
            defer = bool(Fn.FUSE_BNIN and isinstance(nxt, DSConvBNBlock) and nxt.takes_pending_input()
                         and last_bn.track_running_stats and getattr(last_bn, '_tss_sync', None) is None
                         and not observed(module) and Fn.bnin_rows_ok('dw', x))
            if isinstance(module, DSConvBNBlock):
                x = module(x, defer_out=defer)
            else:
                x = module(x, defer_apply=defer)
        return x


class FeatureFusionModule(nn.Module):
    """reference: models/fastscnn.py:67-89.  ``lowres[0]`` stays an ``nn.UpsamplingBilinear2d``
    child (index parity of the state_dict keys ``lowres.1.*`` / ``lowres.2.*``)."""

    def __init__(self, in_channels, out_channels, scale_factor):
        super().__init__()
        lowres_channels, highres_channels = in_channels
        self.scale_factor = scale_factor
        self.lowres = nn.Sequential(
            nn.UpsamplingBilinear2d(scale_factor=scale_factor),
            DWConv2dBlock(lowres_channels, lowres_channels, kernel_size=3,
                          padding=scale_factor, dilation=scale_factor),
            Conv2dBlock(lowres_channels, out_channels, kernel_size=1, use_activation=False),
        )
        self.highres = nn.Sequential(
            Conv2dBlock(highres_channels, out_channels, kernel_size=1, use_activation=False)
        )

    def forward(self, lowres, highres):
        lowres = ops.as_nhwc(lowres)
        s = self.scale_factor
        x = Fn.Bilinear.apply(lowres, lowres.shape[2] * s, lowres.shape[3] * s)
        x = self.lowres[1](x, defer_apply=hands_over_to_pointwise(self.lowres[1], self.lowres[2]))
        high = self.highres[0](highres)
        # relu(lowres + highres): add and ReLU fused into the low-res branch's BatchNorm apply
        return self.lowres[2](x, residual=high, relu=True, sole_consumer=Fn.FUSE_BNRED_EXT)


def Classifier(in_channels, out_channels):
    """reference: models/fastscnn.py:92-98 (also imported by scripts/train_fastscnn.py:28)."""
    head = nn.Sequential(
        DSConv2dBlock(in_channels, in_channels, kernel_size=3, padding=1),
        DSConv2dBlock(in_channels, in_channels, kernel_size=3, padding=1),
        Dropout(0.1),
        ClassScores(in_channels, out_channels),
    )
    head[1].input_sole_consumer = True      # only the second DS block reads the first one's output
    return head


class _GroupedPPM:
    """bins, BatchNorm hyper-parameters and the device address table of the grouped pyramid kernels."""
    __slots__ = ('key', 'bins', 'momentum', 'eps', 'table')


class PyramidPoolingModule(nn.Module):
    """reference: models/fastscnn.py:101-123.  One pooling pass for all bins; the bilinear
    up-samplings write straight into the concat buffer."""

    def __init__(self, in_channels, out_channels, pyramids=(1, 2, 3, 6)):
        super().__init__()
        self.bins = tuple(pyramids)
        self.pyramids = nn.ModuleList([
            nn.Sequential(
                nn.AdaptiveAvgPool2d(bin),
                Conv2dBlock(in_channels, in_channels // len(pyramids), kernel_size=1),
            )
            for bin in pyramids
        ])
        self.conv = Conv2dBlock(in_channels * 2, out_channels, kernel_size=1)

    def _grouped_state(self, x):
        """Device table of addresses for the grouped kernels (csrc/ppm.cu), or None when this call has to take the
        layer-by-layer path (eval mode, no gradient arena, SyncBN, mixed BatchNorm settings)."""
        if not (Fn.FUSE_PPM and self.training and torch.is_grad_enabled()):
            return None
        convs = [p[1][0] for p in self.pyramids]
        bns = [p[1][1] for p in self.pyramids]
        if any(getattr(bn, '_tss_sync', None) is not None or bn.momentum is None or not bn.affine for bn in bns):
            return None
        if len({(float(bn.momentum), float(bn.eps), bn.track_running_stats) for bn in bns}) != 1:
            return None
        if x.dtype != self.pyramids[0][1].compute_dtype or x.shape[0] * min(self.bins) ** 2 < 2:
            return None
        ptrs = []
        for conv, bn in zip(convs, bns):
            grads = [getattr(t, '_tss_grad', None) for t in (conv.weight, bn.weight, bn.bias)]
            if any(g is None for g in grads) or conv.weight.device != x.device:
                return None
            stats = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if bn.track_running_stats else (None,) * 3
            ptrs.append([conv.weight.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr()] +
                        [t.data_ptr() if t is not None else 0 for t in stats] + [g.data_ptr() for g in grads] + [0])
        key = (x.device, tuple(map(tuple, ptrs)))
        st = getattr(self, '_grouped', None)
        if st is None or st.key != key:
            st = self._grouped = _GroupedPPM()
            st.key, st.bins = key, self.bins
            st.momentum, st.eps = float(bns[0].momentum), float(bns[0].eps)
            st.table = torch.tensor(ptrs, dtype=torch.int64).to(x.device)
        return st

    def _grouped_eval(self, x):
        """Eval mode with folded BatchNorm: pool -> all branches -> concat in three launches (instead of ten)."""
        blocks = [p[1] for p in self.pyramids]
        if not (Fn.FUSE_PPM or Fn.FUSE_PPM_EVAL) or self.training or torch.is_grad_enabled() or x.dtype != blocks[0].compute_dtype:
            return None
        if any(not b[1].track_running_stats or b[0].weight.device != x.device for b in blocks):
            return None
        folded = [b._folded(0) for b in blocks]
        ptrs = [[b[0].weight.data_ptr(), f[0].data_ptr(), f[1].data_ptr()] + [0] * 7 for b, f in zip(blocks, folded)]
        key = (x.device, tuple(map(tuple, ptrs)))
        st = getattr(self, '_grouped_ev', None)
        if st is None or st.key != key:
            st = self._grouped_ev = _GroupedPPM()
            st.key, st.bins, st.momentum, st.eps = key, self.bins, 0.0, 0.0
            st.table = torch.tensor(ptrs, dtype=torch.int64).to(x.device)
        N, C, H, W = x.shape
        Cb = blocks[0][0].weight.shape[0]
        pool, _ = ops.adaptive_pool_fwd(x, self.bins)
        z = ops.ppm_branches_eval(pool, st.table, N, C, Cb, self.bins)
        return ops.ppm_concat_fwd(x, z, Cb, self.bins)

    def forward(self, input):
        x = ops.as_nhwc(input)
        cat = self._grouped_eval(x)
        if cat is not None:
            return self.conv(cat)
        st = self._grouped_state(x)
        if st is not None:
            params = [t for p in self.pyramids for t in (p[1][0].weight, p[1][1].weight, p[1][1].bias)]
            return self.conv(Fn.PPMBranches.apply(x, st, *params))
        pooled = Fn.AdaptivePool.apply(x, self.bins)
        zs = [pyramid[1](p) for pyramid, p in zip(self.pyramids, pooled)]
        cat = Fn.PPMConcat.apply(x, *zs)
        return self.conv(cat)


def BottleneckModule(in_channels, out_channels, expansion, repeats=1, stride=1):
    """reference: models/fastscnn.py:126-135."""
    layers = [BottleneckBlock(in_channels, out_channels, expansion=expansion, stride=stride)]
    for _ in range(1, repeats):
        layers.append(BottleneckBlock(out_channels, out_channels, expansion=expansion))
    return nn.Sequential(*layers)
