"""B200 building blocks with the reference's module tree.

The reference builds its blocks as plain ``nn.Sequential(Conv2d, BatchNorm2d[, ReLU])``
(``Conv2dBlock`` fastscnn.py:164-173, ``DWConv2dBlock`` :176-185, ``DSConv2dBlock`` :188-199;
``ConvBlock``/``DWConvBlock`` contextnet.py:150-177).  The classes below ARE such
``nn.Sequential``s -- same children, same indices, same parameter shapes and the same
construction order (identical ``state_dict`` keys and identical random init under the same
seed) -- but their ``forward`` runs the fused sm_100a kernels instead of the children's
forwards.  The child modules only hold parameters and buffers.
"""
import torch
from torch import nn

from .. import ops
from .. import functional as Fn


import os

# inference only: depthwise 3x3 + the pointwise conv that follows it as one kernel (csrc/dwpw_tc.cu).
# One CTA owns an 8 x 16 pixel tile and walks ALL input channels, so the fused kernel needs enough tiles
# to fill the GPU: measured on B200 at 1024x2048, batch 16: 4170 -> 4358 img/s, batch 1 (16..64 tiles per
# layer): 2179 -> 1838 FPS.  Below FUSE_MIN_TILES the two-kernel path is used.
FUSE_DW_PW = os.environ.get('TSS_FUSE_DWPW', '1') == '1'
FUSE_MIN_TILES = int(os.environ.get('TSS_FUSE_MIN_TILES', '148'))     # one tile per SM; see the measurement above


def fused_dw_pw(dw_block, dw_ci, pw_block, pw_ci, x, relu1, relu2, residual=None):
    """Eval-mode ``dw3x3 -> BN [-> ReLU] -> 1x1 -> BN [+ residual] [-> ReLU]`` through ``tss_dwpw_fwd``, or
    None when the shapes are not covered (stride / dilation 1, channels % 64, bf16 tensor-core mode)."""
    dw, pw = dw_block[dw_ci], pw_block[pw_ci]
    if not FUSE_DW_PW or dw_block.training or pw_block.training or torch.is_grad_enabled():
        return None
    if dw.stride != (1, 1) or dw.dilation != (1, 1) or dw.in_channels % 64 or pw.out_channels % 16 \
            or pw.out_channels > 256 or dw_block.compute_dtype != torch.bfloat16:
        return None
    tiles = x.shape[0] * ((x.shape[2] + 7) // 8) * ((x.shape[3] + 15) // 16)
    if tiles < FUSE_MIN_TILES:
        return None
    packed = pw_block._packed(pw_ci)
    if packed is None:
        return None
    x = ops.as_nhwc(x)
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    s1, b1 = dw_block._folded(dw_ci)
    s2, b2 = pw_block._folded(pw_ci)
    return ops.dwpw_fwd(x, dw.weight, s1, b1, relu1, packed[0], s2, b2,
                        res=ops.as_nhwc(residual) if residual is not None else None, relu2=relu2)


def _versions(*tensors):
    return (ops.WEIGHTS_EPOCH[0],) + tuple(t._version for t in tensors) + tuple(t.data_ptr() for t in tensors)


class _FusedConvBN:
    """Mixin state/logic shared by all conv+BN pairs (conv at index ``ci``, BN at ``ci+1``)."""

    def _init_fused(self):
        self._specs = {}
        self._fold_cache = {}
        self._pack_cache = {}
        self.compute_dtype = torch.float32
        self.pw_impl = 0

    def _kind(self, conv):
        if conv.kernel_size == (1, 1) and conv.groups == 1:
            return 'pw'
        if conv.kernel_size == (3, 3) and conv.groups == conv.in_channels and conv.in_channels == conv.out_channels \
                and conv.padding == conv.dilation and conv.dilation[0] == conv.dilation[1]:
            return 'dw'
        if conv.kernel_size == (3, 3) and conv.groups == 1 and conv.in_channels == 3 and conv.stride == (2, 2) \
                and conv.padding == (1, 1):
            return 'stem'
        if conv.kernel_size == (3, 3) and conv.groups == 1 and conv.in_channels % 8 == 0 and conv.stride == (1, 1) \
                and conv.padding == (1, 1) and conv.dilation == (1, 1):
            return 'dense3'
        raise RuntimeError('no sm_100a kernel for %r' % (conv,))

    def _spec(self, ci, relu):
        key = (ci, relu, self.compute_dtype, self.pw_impl)
        spec = self._specs.get(key)
        if spec is None:
            conv, bn = self[ci], self[ci + 1]
            kind = self._kind(conv)
            # a dense 3x3 runs as the pointwise GEMM over its patch matrix (csrc/conv3x3.cu)
            spec = Fn.ConvSpec('pw' if kind == 'dense3' else kind, conv.stride[0], conv.dilation[0], relu, bn,
                               self.compute_dtype, self.pw_impl)
            self._specs[key] = spec
        return spec

    def _folded(self, ci):
        """Eval-mode BatchNorm as per-channel scale/shift, cached on the tensors' versions."""
        bn = self[ci + 1]
        key = _versions(bn.weight, bn.bias, bn.running_mean, bn.running_var)
        hit = self._fold_cache.get(ci)
        if hit is None or hit[0] != key:
            # refreshed IN PLACE once it exists: captured eval graphs and address tables keep reading the same buffers
            pair = ops.bn_fold(bn, out=hit[1][0]._base if (hit is not None and hit[1][0]._base is not None) else None)
            hit = (key, pair)
            self._fold_cache[ci] = hit
        return hit[1]

    def _packed(self, ci):
        """bf16 (W, W^T) copies for the tensor-core path, cached on the weight's version."""
        if self.pw_impl == 0 or self.compute_dtype != torch.bfloat16:
            return None
        w = self[ci].weight
        if self._kind(self[ci]) != 'pw' or w.shape[0] % 16 != 0 or w.shape[1] % 16 != 0:
            return None
        opt = getattr(w, '_tss_optimizer', None)
        if opt is not None:
            # weight lives in a FlatAdamW arena: persistent packs refreshed by the optimizer's single
            # multi-tensor launch after each step; only an outside in-place write (load_state_dict)
            # bumps the tensor version and triggers a re-pack here
            key = (id(opt), w._version, w.data_ptr())
            hit = self._pack_cache.get(ci)
            if hit is None or hit[0] != key:
                hit = (key, opt.register_pack(w))
                self._pack_cache[ci] = hit
            return hit[1]
        key = _versions(w)
        hit = self._pack_cache.get(ci)
        if hit is None or hit[0] != key:
            hit = (key, ops.pack_weights_bf16(w, out=hit[1] if hit is not None else None))      # in place, as above
            self._pack_cache[ci] = hit
        return hit[1]

    def _conv_bn(self, ci, x, relu, res=None, sole_consumer=False, defer_apply=False):
        """``sole_consumer``: ``x`` is the output of another fused conv+BN pair and nothing else reads it (the
        enclosing block guarantees it), so that layer's BatchNorm-backward reduction may be folded into this
        conv's dgrad epilogue (functional.FUSE_BNRED)."""
        producer = getattr(x, '_tss_bn_link', None) if sole_consumer else None
        in_affine = getattr(x, '_tss_in_affine', None)       # x is a RAW conv output whose BatchNorm this conv applies
        conv, bn = self[ci], self[ci + 1]
        spec = self._spec(ci, relu)
        if spec.kind != 'stem':
            x = ops.as_nhwc(x)
            if x.dtype != self.compute_dtype:
                x = x.to(self.compute_dtype)
        if res is not None:
            res = ops.as_nhwc(res)
        weight = conv.weight
        if self._kind(conv) == 'dense3':
            x = Fn.Im2Col3x3.apply(x)
            weight = Fn.TapMajorWeight.apply(weight)
            packed = None
            if self.pw_impl != 0 and self.compute_dtype == torch.bfloat16 and weight.shape[0] % 16 == 0 \
                    and weight.shape[1] % 16 == 0:
                packed = ops.pack_weights_bf16(weight.detach())
        else:
            packed = self._packed(ci)
        use_batch_stats = self.training or not bn.track_running_stats
        if use_batch_stats:
            # `defer_apply`: the enclosing block promises that the single consumer is a depthwise conv that applies
            # this layer's BatchNorm + ReLU itself (functional.FUSE_BNIN); `_tss_in_affine` is that hand-over
            if in_affine is not None and not ((spec.kind == 'dw' and spec.dilation == 1 and spec.stride in (1, 2))
                                              or (spec.kind == 'pw' and packed is not None and spec.impl == 1)):
                raise RuntimeError('a tensor with a pending BatchNorm reached a layer that cannot apply it')
            z = Fn.ConvBNAct.apply(x, res, weight, bn.weight, bn.bias, spec, packed, producer, defer_apply, in_affine)
            if Fn.ConvBNAct.last_link is not None:
                z._tss_bn_link, Fn.ConvBNAct.last_link = Fn.ConvBNAct.last_link, None
            if Fn.ConvBNAct.last_affine is not None:
                z._tss_in_affine, Fn.ConvBNAct.last_affine = Fn.ConvBNAct.last_affine, None
            return z
        if in_affine is not None:
            raise RuntimeError('a tensor with a pending BatchNorm reached an eval-mode layer')
        if torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
            raise RuntimeError('eval-mode BatchNorm with autograd is not implemented; '
                               'wrap inference in torch.no_grad()')
        scale, shift = self._folded(ci)
        return Fn.conv_forward(spec, x, weight, scale=scale, shift=shift, res=res, relu=relu,
                               packed=packed)


class ConvBNBlock(nn.Sequential, _FusedConvBN):
    """``Conv2dBlock`` / ``DWConv2dBlock`` (fastscnn.py:164-185), ``ConvBlock`` / ``DWConvBlock``
    (contextnet.py:150-177): Conv2d(bias=False) -> BatchNorm2d -> optional ReLU(inplace)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, use_activation=True):
        layers = [
            nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                      dilation=dilation, groups=groups, bias=False),
            nn.BatchNorm2d(out_channels),
        ]
        if use_activation:
            layers.append(nn.ReLU(inplace=True))
        super().__init__(*layers)
        self.use_activation = use_activation
        self._init_fused()

    def forward(self, input, residual=None, relu=None, sole_consumer=False, defer_apply=False):
        """``residual``/``relu`` let the enclosing block fuse its ``+input`` and trailing
        ``F.relu`` (fastscnn.py:158-161, 89) into this block's BatchNorm apply."""
        return self._conv_bn(0, input, self.use_activation if relu is None else relu, residual, sole_consumer, defer_apply)


class DSConvBNBlock(nn.Sequential, _FusedConvBN):
    """``DSConv2dBlock`` fastscnn.py:188-199: dw3x3 -> BN (no ReLU) -> 1x1 -> BN -> optional ReLU."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 use_activation=True):
        layers = [
            nn.Conv2d(in_channels, in_channels, kernel_size, stride=stride, padding=padding,
                      dilation=dilation, groups=in_channels, bias=False),
            nn.BatchNorm2d(in_channels),
            nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels),
        ]
        if use_activation:
            layers.append(nn.ReLU(inplace=True))
        super().__init__(*layers)
        self.use_activation = use_activation
        self._init_fused()

    def takes_pending_input(self):
        """This block's depthwise half can apply a producer's BatchNorm + ReLU while reading (functional.FUSE_BNIN)."""
        c = self[0]
        return bool(self.training and torch.is_grad_enabled() and c.dilation[0] == 1 and c.stride[0] in (1, 2)
                    and (c.in_channels % 32 == 0 or c.in_channels % 48 == 0) and getattr(self[1], '_tss_sync', None) is None)

    def forward(self, input, defer_out=False):
        """``defer_out``: the enclosing chain promises that the only reader of this block's output is a depthwise conv
        that applies the pointwise half's BatchNorm + ReLU itself."""
        y = fused_dw_pw(self, 0, self, 2, input, False, self.use_activation)
        if y is not None:
            return y
        # `input_sole_consumer` (set by the enclosing model where it holds): nothing but this block reads `input`
        # the pointwise half can apply the depthwise half's BatchNorm in its operand producer (functional.FUSE_BNIN_PW)
        defer = bool(Fn.FUSE_BNIN_PW and self.training and torch.is_grad_enabled() and self.pw_impl == 1
                     and self.compute_dtype == torch.bfloat16 and self[2].in_channels % 16 == 0 and self[2].out_channels % 16 == 0
                     and self[1].track_running_stats and getattr(self[1], '_tss_sync', None) is None
                     and getattr(self[3], '_tss_sync', None) is None)
        x = self._conv_bn(0, input, False, sole_consumer=Fn.FUSE_BNRED_EXT and getattr(self, 'input_sole_consumer', False),
                          defer_apply=defer)
        return self._conv_bn(2, x, self.use_activation, sole_consumer=True, defer_apply=defer_out)


def observed(module):
    """Somebody looks at this module's output through a forward hook (deep-supervision taps, feature extraction): its
    BatchNorm apply pass must then not be handed over to the consumer, or the hook would see the raw convolution."""
    from torch.nn.modules import module as _m
    return bool(module._forward_hooks or _m._global_forward_hooks)


def hands_over_to_pointwise(dw, pw):
    """The depthwise block ``dw`` may leave its BatchNorm (+ReLU) to the tensor-core pointwise block ``pw`` that is its
    only reader (functional.FUSE_BNIN_PW)."""
    if not (Fn.FUSE_BNIN_PW and isinstance(dw, ConvBNBlock) and isinstance(pw, ConvBNBlock) and dw.training
            and torch.is_grad_enabled() and not observed(dw)):
        return False
    c_dw, c_pw = dw[0], pw[0]
    return bool(c_dw.groups == c_dw.in_channels and c_dw.groups > 1 and c_pw.kernel_size == (1, 1) and c_pw.groups == 1
                and pw.pw_impl == 1 and pw.compute_dtype == torch.bfloat16 and c_pw.in_channels % 16 == 0
                and c_pw.out_channels % 16 == 0 and dw[1].track_running_stats
                and getattr(dw[1], '_tss_sync', None) is None and getattr(pw[1], '_tss_sync', None) is None)


def Conv2dBlock(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                use_activation=True):
    return ConvBNBlock(in_channels, out_channels, kernel_size, stride, padding, dilation, 1, use_activation)


def DWConv2dBlock(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                  use_activation=True):
    if in_channels != out_channels:
        raise ValueError("input and output channels must be the same in depthwise convolution")
    return ConvBNBlock(in_channels, out_channels, kernel_size, stride, padding, dilation, in_channels,
                       use_activation)


def DSConv2dBlock(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                  use_activation=True):
    return DSConvBNBlock(in_channels, out_channels, kernel_size, stride, padding, dilation, use_activation)


class BottleneckBlock(nn.Module):
    """``BottleneckBlock`` fastscnn.py:138-161 / contextnet.py:129-147.  The residual add and the
    unconditional trailing ReLU are fused into conv3's BatchNorm apply."""

    def __init__(self, in_channels, out_channels, stride=1, expansion=6):
        super().__init__()
        expansion_channels = expansion * in_channels
        self.conv1 = Conv2dBlock(in_channels, expansion_channels, kernel_size=1)
        self.conv2 = DWConv2dBlock(expansion_channels, expansion_channels, kernel_size=3, padding=1,
                                   stride=stride)
        self.conv3 = Conv2dBlock(expansion_channels, out_channels, kernel_size=1, use_activation=False)
        self.has_residual = stride == 1 and in_channels == out_channels   # "x.shape == input.shape"

    def _defer_conv1_apply(self):
        """conv2 can apply conv1's BatchNorm + ReLU while reading its input (functional.FUSE_BNIN)."""
        c2, bn1 = self.conv2[0], self.conv1[1]
        return bool(Fn.FUSE_BNIN and self.training and torch.is_grad_enabled() and c2.dilation[0] == 1
                    and c2.stride[0] in (1, 2) and c2.in_channels % 32 == 0 and bn1.track_running_stats
                    and getattr(bn1, '_tss_sync', None) is None and getattr(self.conv2[1], '_tss_sync', None) is None)

    def _defer_conv2_apply(self):
        """conv3 (tensor-core pointwise) can apply conv2's BatchNorm + ReLU in its operand producer
        (functional.FUSE_BNIN_PW)."""
        c3 = self.conv3
        return bool(Fn.FUSE_BNIN_PW and self.training and torch.is_grad_enabled() and c3.pw_impl == 1
                    and c3.compute_dtype == torch.bfloat16 and c3[0].in_channels % 16 == 0 and c3[0].out_channels % 16 == 0
                    and self.conv2[1].track_running_stats and getattr(self.conv2[1], '_tss_sync', None) is None
                    and getattr(c3[1], '_tss_sync', None) is None)

    def forward(self, input):
        x = self.conv1(input, defer_apply=self._defer_conv1_apply() and not observed(self.conv1) and Fn.bnin_rows_ok('dw', input))
        res = input if self.has_residual else None
        y = fused_dw_pw(self.conv2, 0, self.conv3, 0, x, self.conv2.use_activation, True, residual=res)
        if y is not None:
            return y
        x = self.conv2(x, sole_consumer=True,
                       defer_apply=self._defer_conv2_apply() and not observed(self.conv2) and Fn.bnin_rows_ok('pw', input))
        return self.conv3(x, residual=res, relu=True, sole_consumer=True)


class ClassScores(nn.Conv2d):
    """``nn.Conv2d(in_channels, classes, 1)`` with bias (fastscnn.py:97, contextnet.py:86)."""

    def __init__(self, in_channels, out_channels):
        super().__init__(in_channels, out_channels, kernel_size=1)
        self.compute_dtype = torch.float32

    def forward(self, input):
        x = ops.as_nhwc(input)
        if x.dtype != self.compute_dtype:
            x = x.to(self.compute_dtype)
        return Fn.ConvBias.apply(x, self.weight, self.bias)


def convert_syncbn_model(module, process_group=None):
    """Counterpart of ``apex.parallel.convert_syncbn_model`` (scripts/train_fastscnn.py:145): every
    BatchNorm2d of the fused blocks normalises with the statistics of ALL ranks' batches (one fp64
    all-reduce of the 2C channel sums per layer in forward, one fp32 all-reduce of the 2C gradient
    sums per layer in backward).  Every rank must hold the same number of pixels per step.  The
    module is modified in place and returned.  Per-rank statistics (the default) avoid these 88
    latency-bound collectives per step; see DESIGN.md section 6."""
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d):
            m._tss_sync = True if process_group is None else process_group
    return module


class Dropout(nn.Dropout):
    """``nn.Dropout`` (same constructor, no state) whose training forward can run on the library's kernel
    (``functional.OWN_DROPOUT``); everything else -- eval mode, p = 0, CPU tensors -- is the stock module."""

    def forward(self, input):
        if (Fn.OWN_DROPOUT and self.training and 0.0 < self.p < 1.0 and input.dim() == 4 and input.shape[1] % 8 == 0
                and ops.geom(input) is not None and ops.geom(input)[4] == input.shape[1]):
            return Fn.Dropout.apply(input, self.p)
        return super().forward(input)


def refresh_cached_operands(module):
    """Bring every derived eval-mode operand that has been built so far (folded BatchNorm scale/shift, bf16 weight packs)
    up to date IN PLACE.  A captured CUDA graph does not run the Python that would notice a stale cache, so whoever
    replays eval graphs calls this first (engine.create_segmentation_evaluator does, once per run)."""
    for m in module.modules():
        if isinstance(m, _FusedConvBN):
            for ci in list(m._fold_cache):
                m._folded(ci)
            for ci in list(m._pack_cache):
                m._packed(ci)


def set_compute_dtype(module, dtype, pw_impl=None):
    """Select the activation storage type (fp32 verification mode / bf16) of every fused block."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError('compute dtype must be float32 or bfloat16')
    for m in module.modules():
        if hasattr(m, 'compute_dtype'):
            m.compute_dtype = dtype
        if pw_impl is not None and hasattr(m, 'pw_impl'):
            m.pw_impl = pw_impl
    return module
