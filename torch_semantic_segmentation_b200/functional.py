"""Autograd formulas of the hot path, expressed with the C-ABI kernels only.

One ``torch.autograd.Function`` per reference building block:

* :class:`ConvBNAct`   -- Conv2d(bias=False) -> BatchNorm2d -> [+residual] -> [ReLU]
  (``Conv2dBlock``/``DWConv2dBlock`` fastscnn.py:164-185, the tail of ``BottleneckBlock``
  fastscnn.py:158-161 and of ``FeatureFusionModule`` fastscnn.py:89), in training mode:
  conv kernel (+BatchNorm statistics in its epilogue) -> finalize -> apply.
* :class:`ConvBias`    -- the classifier's Conv2d(128, classes, 1) with bias (fastscnn.py:97).
* :class:`AdaptivePool`, :class:`PPMConcat` -- ``PyramidPoolingModule`` fastscnn.py:108,119-122.
* :class:`Bilinear`, :class:`UpsampleLogits` -- ``align_corners=True`` resizes
  (fastscnn.py:74, 63-64).
* :class:`CrossEntropy` -- ``nn.CrossEntropyLoss(ignore_index=255)`` (train_fastscnn.py:132),
  forward and gradient fused in one pass.

Backward runs on autograd's worker thread; every kernel launch goes to that thread's
current stream (``_lib.call``), which autograd sets to the forward stream.
"""
import os
import threading

import torch

from . import library  # noqa: F401  (registers the tss_b200 operators)
from . import ops
from .gates import gate

_tls = threading.local()


class unit_loss_grad:
    """``with unit_loss_grad(loss): loss.backward()`` -- tells the loss Function that produced ``loss`` that the
    gradient it is about to receive is the implicit seed 1.0, so that it can hand its fused gradient on without a
    rescaling pass.  The mark is put on ``loss.grad_fn`` itself and only if that node IS one of this module's loss
    Functions: a loss term that reaches the total through any other node (``0.4 * ce(aux)``, a sum of terms) is not
    marked and multiplies its gradient by the ``grad_out`` autograd hands it (train_fastscnn.py:134-136 weights
    the auxiliary terms by 0.4)."""

    def __init__(self, loss=None):
        node = getattr(loss, 'grad_fn', None)
        self.node = node if (node is not None and type(node).__name__ in _SEEDED_LOSS_NODES) else None

    def __enter__(self):
        if self.node is not None:
            self.node._tss_unit_seed = True
        return self

    def __exit__(self, *exc):
        if self.node is not None:
            self.node._tss_unit_seed = False


# autograd names the backward node of ``class X(Function)`` ``XBackward``
_SEEDED_LOSS_NODES = ('UpsampleCrossEntropyBackward', 'OhemCrossEntropyBackward', 'CrossEntropyBackward')


def _unit_grad(ctx):
    return getattr(ctx, '_tss_unit_seed', False)


_grad_ready_hook = None


def set_grad_ready_hook(fn):
    """``fn(param)`` is called from backward right after the kernels that produce ``param``'s
    gradient have been enqueued (used by ``distributed.GradientAllReducer`` to launch bucket
    all-reduces while the rest of backward is still running)."""
    global _grad_ready_hook
    _grad_ready_hook = fn


def grad_ready(*params):
    if _grad_ready_hook is not None:
        for p in params:
            if p is not None:
                _grad_ready_hook(p)


class _WgradLane:
    """Weight-gradient kernels run on a side stream: in backward nothing on the critical chain
    (BatchNorm backward -> dgrad -> next layer) depends on a wgrad, so the ~45 wgrad launches of a
    step overlap with it instead of lengthening it.  Only used when the gradients are accumulated
    in place into the optimizer's arena (``FlatAdamW``), whose ``step()`` / all-reduce joins the lane.
    Operand tensors are kept referenced until the join so that the caching allocator cannot hand
    their memory to a main-stream kernel while a side-stream wgrad still reads it."""

    def __init__(self):
        self.enabled = os.environ.get('TSS_WGRAD_LANE', '1') != '0'
        self.streams = {}
        self.keep = []
        self.dirty = set()
        self.queued = False

    def stream(self, device):
        st = self.streams.get(device)
        if st is None:
            st = self.streams[device] = torch.cuda.Stream(device)
        return st

    def run(self, device, fn, *tensors):
        if not (self.enabled and device.type == 'cuda'):
            fn()
            return
        side = self.stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            fn()
        self.keep.append(tensors)
        self.dirty.add(device)
        if not self.queued:       # join when this backward pass ends, so that p.grad is complete for
            self.queued = True    # whatever the caller does next on its stream (clipping, step, ...)
            torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)

    def _end_of_backward(self):
        self.queued = False
        self.join()

    def join(self, device=None):
        """Make the current stream wait for every wgrad launched so far."""
        for dev in list(self.dirty):
            if device is None or dev == device:
                torch.cuda.current_stream(dev).wait_stream(self.streams[dev])
                self.dirty.discard(dev)
        if not self.dirty:
            self.keep.clear()


wgrad_lane = _WgradLane()


class ConvSpec:
    """Static description of one conv+BN block (built once per module)."""

    __slots__ = ('kind', 'stride', 'dilation', 'relu', 'bn', 'dtype', 'impl')

    def __init__(self, kind, stride, dilation, relu, bn, dtype, impl=0):
        self.kind, self.stride, self.dilation = kind, stride, dilation
        self.relu, self.bn, self.dtype, self.impl = relu, bn, dtype, impl


def conv_forward(spec, x, weight, scale=None, shift=None, res=None, relu=False, stats=None, packed=None):
    """The raw convolution of a block with an optional fused epilogue."""
    if spec.kind == 'pw':
        wp = packed[0] if packed is not None else None
        return ops.pwconv_fwd(x, weight, scale=scale, shift=shift, res=res, relu=relu, stats=stats,
                              wp=wp, impl=spec.impl if wp is not None else 0)
    if res is not None:
        raise RuntimeError('residual epilogue is only available on pointwise convolutions')
    if spec.kind == 'dw':
        return ops.dwconv_fwd(x, weight, spec.stride, spec.dilation, scale=scale, shift=shift,
                              relu=relu, stats=stats)
    if spec.kind == 'stem':
        if STEM_TC and spec.dtype == torch.bfloat16 and weight.shape[0] == 32:
            return ops.stem_fwd_tc(x, weight, scale=scale, shift=shift, relu=relu, stats=stats)
        return ops.stem_fwd(x, weight, spec.dtype, scale=scale, shift=shift, relu=relu, stats=stats)
    raise RuntimeError('unknown conv kind %r' % spec.kind)


# Stem convolution on tcgen05 with a thread-built implicit-GEMM operand (csrc/stem_tc.cu) instead of the SIMT
# kernel (864 FMAs per output pixel).  On by default (B200: forward 154 -> 95 us).
STEM_TC = gate('STEM_TC')
# With the tensor-core stem: its BatchNorm-backward apply inside the weight gradient's operand producer -- the largest
# activation of the net (32 channels at 1/2 resolution) is not rewritten as dy.  Validated on the B200, slower in the step
# (4.71 vs 4.52 ms): off unless TSS_STEM_BWD_FUSED=1.
STEM_BWD_FUSED = gate('STEM_BWD_FUSED')

# Training: fold the BatchNorm-backward reduction of a producer layer into the dgrad epilogue of its single
# consumer (csrc/pwconv_tc_bnred.cu, csrc/dwconv_bnred.cu): one full read of (dz, y) and one launch less per
# fused layer (20 of the 44 BatchNorm layers of Fast-SCNN; 4.66 -> 4.57 ms/step on B200).  TSS_FUSE_BNRED=0
# selects the stand-alone reduction everywhere.
FUSE_BNRED = gate('FUSE_BNRED')
# Extended set: the stride-2 depthwise dgrad (its producers are the stem and the first expand conv: the largest
# BatchNorm-backward instances) and two more single-consumer pairs at 1/8 resolution (fusion low-res branch,
# classifier).  On by default (B200: 4.58 -> 4.48 ms/step).
FUSE_BNRED_EXT = gate('FUSE_BNRED_EXT')
# BatchNorm-backward APPLY folded into the A-operand producer of the pointwise dgrad (csrc/pwconv_tc_bwd.cu):
# dy is formed in registers and goes straight into the swizzled shared-memory tile of the tcgen05 GEMM (one
# launch and one read of dy less per 1x1 layer without a residual).  Validated on the B200, slower in the step (4.64 vs
# 4.58 ms: thread-built operands lose against TMA-fed ones on the small maps): off unless TSS_FUSE_BNAPPLY=1.
FUSE_BNAPPLY = gate('FUSE_BNAPPLY')
# The same for the stride-1 depthwise layers whose dgrad already carries the producer's reduction
# (csrc/dwconv_bwd_fused.cu: dz and y arrive as two TMA halo tiles, dy replaces dz in shared memory).
# Validated on the B200, no gain in the step: off unless TSS_FUSE_BNAPPLY_DW=1.
FUSE_BNAPPLY_DW = gate('FUSE_BNAPPLY_DW')
# The four pyramid-pooling branches as grouped launches (csrc/ppm.cu): 3 launches forward and 5 backward instead
# of ~18 and ~26.  Validated on the B200 (except 2 samples per channel in fp32, an open tolerance item), no gain in the
# training step: off unless TSS_FUSE_PPM=1.
FUSE_PPM = gate('FUSE_PPM')
# ... in eval mode only (folded BatchNorm): 7 launches less per inference forward (bs1 1024x2048: 2366 -> 2425 FPS on B200)
FUSE_PPM_EVAL = gate('FUSE_PPM_EVAL')
# BatchNorm finalize folded into the apply kernel (csrc/bn_fused.cu): one launch less per layer on the forward
# chain (44 per step).  Validated on the B200 (the whole GPU suite passes with the gate on), still slower in the step with
# the constants derived once per CTA (3.204 vs 3.185 ms; per thread: 4.19 vs 3.80): off unless TSS_FUSE_BNFIN=1.
FUSE_BNFIN = gate('FUSE_BNFIN')
# Inside a bottleneck, conv1's BatchNorm + ReLU applied by conv2 (depthwise) while it reads its input tile
# (csrc/dwconv_bnin.cu): the expanded activation is never materialised, conv1's apply pass disappears.
# Kernels validated on the B200, on by default since the reference cycle that broke graph capture is gone (3.67 -> 3.58 ms/step).
FUSE_BNIN = gate('FUSE_BNIN')
# The same hand-over from conv2 (depthwise) to conv3 (tensor-core pointwise, csrc/pwconv_tc_fwd_bnin.cu): the activated
# tensor is written once by the GEMM's operand producer (the weight gradient needs it) and never read in the forward
# pass.  Same status as FUSE_BNIN: off unless TSS_FUSE_BNIN_PW=1.
FUSE_BNIN_PW = gate('FUSE_BNIN_PW')


def _row_range(name):
    lo, _, hi = os.environ.get(name, ':').partition(':')
    return (int(lo) if lo else 0, int(hi) if hi else 1 << 62)


_BNIN_ROWS = {'dw': _row_range('TSS_BNIN_ROWS'), 'pw': _row_range('TSS_BNIN_PW_ROWS')}


def bnin_rows_ok(kind, t):
    """A/B aid: ``TSS_BNIN_ROWS=lo:hi`` / ``TSS_BNIN_PW_ROWS=lo:hi`` restrict the hand-over of a BatchNorm apply pass to
    its depthwise / pointwise consumer to tensors of lo <= N*H*W <= hi pixels (``t``: the NCHW-shaped tensor handed over)."""
    lo, hi = _BNIN_ROWS[kind]
    rows = t.shape[0] * t.shape[2] * t.shape[3]
    return lo <= rows <= hi


class _BnLink:
    """What the consumer's dgrad needs to run the producer's BatchNorm-backward reduction: the producer's raw
    conv output ``y``, its batch statistics and affine parameters, whether a ReLU follows, and the (zeroed)
    ``sums`` buffer of the producer's per-layer scratch.  ``reduced`` tells the producer's backward that the
    gradient it receives already is ``g = dz * mask`` and that ``sums`` is complete."""

    __slots__ = ('y', 'mean', 'rstd', 'gamma', 'beta', 'relu', 'sums', 'bn', 'scratch', 'reduced')

    def __init__(self, y, mean, rstd, gamma, beta, relu, scratch, bn, C):
        self.y, self.mean, self.rstd, self.gamma, self.beta, self.relu = y, mean, rstd, gamma, beta, relu
        self.scratch, self.bn = scratch, bn
        self.sums = scratch[2 * C:].view(torch.float32)
        self.reduced = False

    def usable(self):
        """The producer's sums buffer is still the zeroed one of its last forward."""
        return not self.reduced and not self.bn._tss_dirty and self.scratch is getattr(self.bn, '_tss_scratch', None)


def _sync_group(bn):
    """-> (world size, process group) if this BatchNorm layer synchronises its statistics across ranks."""
    sync = getattr(bn, '_tss_sync', None)
    if sync is None or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return 1, None
    group = None if sync is True else sync
    return torch.distributed.get_world_size(group), group


def _finished_sums(ctx, dz):
    """The two BatchNorm-backward sums of the layer, for the kernels that fold its apply pass into something else:
    -> (sums, mask).  Either the consumer's dgrad epilogue already produced them (``dz`` is then masked: mask False), or
    the stand-alone reduction runs now (mask = the layer's ReLU, recomputed from the raw output by the fused kernel)."""
    spec, link = ctx.spec, ctx.link
    x, weight, gamma, beta, y, z, mean, rstd = ctx.saved_tensors
    C = weight.shape[0]
    if link is not None and link.reduced:
        return link.sums, False
    sums = ctx.scratch[2 * C:].view(torch.float32)
    if spec.bn._tss_dirty or ctx.scratch is not getattr(spec.bn, '_tss_scratch', None):
        sums = ops.zeros_f32(2 * C, weight.device)      # a second backward without a forward in between
    spec.bn._tss_dirty = True
    ops.bn_backward_reduce(dz, y, mean, rstd, gamma, beta, spec.relu, sums)
    return sums, spec.relu


def _conv_grads(ctx, dx, dres, dw, gg_out, gb_out):
    """What ConvBNAct.backward returns: gradients that were accumulated straight into the optimizer's arena are
    reported as None (and announced to the reducer through grad_ready)."""
    gw, gg, gb = ctx.arena
    grad_ready(*ctx.params)
    return (dx, dres, None if gw is not None else dw, None if gg is not None else gg_out,
            None if gb is not None else gb_out, None, None, None, None, None)


def _dw_wgrad_fn(ctx, x, dy, dw, spec):
    """The depthwise weight gradient of a layer: on the activated input, or on the producer's raw output with its
    BatchNorm applied on the fly when the forward pass ran that way."""
    aff = ctx.in_affine
    if aff is not None:
        return lambda: ops.dwconv_wgrad_bnin(x, aff[0], aff[1], aff[2], dy, dw, spec.stride)
    return lambda: ops.dwconv_wgrad(x, dy, dw, spec.stride, spec.dilation)


class ConvBNAct(torch.autograd.Function):
    """z = act(BN_train(conv(x, w)) [+ res])."""

    last_link = None        # handed to nn.blocks right after apply(): the link of the layer just run
    last_affine = None      # (scale, shift, relu) of a layer whose apply pass was left to its consumer

    @staticmethod
    def forward(ctx, x, res, weight, gamma, beta, spec, packed, producer=None, defer_apply=False, in_affine=None):
        bn = spec.bn
        C = weight.shape[0]
        if bn.momentum is None:
            raise RuntimeError('BatchNorm2d(momentum=None) (cumulative average) is not supported')
        # per-layer scratch [statistics 2C fp64 | backward sums 2C fp32]: finalize re-zeroes all of it
        scratch = ops.layer_scratch(bn, weight.device)
        x_saved = x
        if in_affine is not None and spec.kind == 'pw':
            # x is the RAW output of the producer: the GEMM's operand producer applies its BatchNorm (+ReLU) and stores
            # the activated tensor once -- that is what the weight gradient (and nothing else) reads
            y, x_saved = ops.pwconv_fwd_bnin(x, in_affine[0], in_affine[1], in_affine[2], packed[0], scratch)
            in_affine = None
        elif in_affine is not None:    # depthwise: applied behind every read of the input tile, forward and wgrad
            y = ops.dwconv_fwd_bnin(x, in_affine[0], in_affine[1], in_affine[2], weight, spec.stride, scratch)
        else:
            y = conv_forward(spec, x, weight, stats=scratch, packed=packed)
        ctx.in_affine = in_affine
        ctx.patches = None
        if (spec.kind == 'stem' and STEM_TC and gate('STEM_WGRAD_PATCHES') and spec.dtype == torch.bfloat16 and C == 32
                and x.is_cuda and wgrad_lane.enabled and getattr(weight, '_tss_grad', None) is not None):
            # the stem's weight gradient reads the image as a patch matrix that does not depend on the backward pass:
            # built now, on the weight-gradient side stream, under the rest of the forward pass
            side = wgrad_lane.stream(x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                ctx.patches = ops.stem_patches(x)
        N, _, H, W = y.shape
        world, group = _sync_group(bn)
        if world > 1:
            # SyncBN (apex convert_syncbn_model, scripts/train_fastscnn.py:145): the per-channel sums of
            # all ranks, one fp64 all-reduce per layer; every rank holds the same number of pixels
            torch.distributed.all_reduce(scratch[:2 * C], group=group)
        ctx.sync = (world, group)
        ConvBNAct.last_affine = None
        if defer_apply:
            # the single consumer (a depthwise conv, nn.blocks) applies scale / shift / ReLU while reading: z is y
            scale, shift, mean, rstd = ops.bn_finalize(scratch, N * H * W * world, bn, float(bn.momentum), float(bn.eps),
                                                       update_running=bn.track_running_stats, clear_n=3 * C, C=C)
            z = y
            ConvBNAct.last_affine = (scale, shift, spec.relu)
        elif FUSE_BNFIN:
            z, mean, rstd = ops.bn_finalize_apply(scratch, N * H * W * world, bn, float(bn.momentum), float(bn.eps), y, res=res,
                                                  relu=spec.relu, update_running=bn.track_running_stats, clear_n=3 * C, C=C)
        else:
            scale, shift, mean, rstd = ops.bn_finalize(scratch, N * H * W * world, bn, float(bn.momentum), float(bn.eps),
                                                       update_running=bn.track_running_stats, clear_n=3 * C, C=C)
            z = ops.bn_apply(y, scale, shift, res=res, relu=spec.relu)
        bn._tss_dirty = False
        ctx.scratch = scratch
        ctx.spec, ctx.packed = spec, packed
        ctx.params = (weight, gamma, beta)
        ctx.arena = (getattr(weight, '_tss_grad', None), getattr(gamma, '_tss_grad', None),
                     getattr(beta, '_tss_grad', None))
        ctx.has_res = res is not None
        ctx.in_hw = (x.shape[2], x.shape[3])
        # the ReLU mask is recomputed from y in backward unless a residual was added before the ReLU
        ctx.save_for_backward(x_saved, weight, gamma, beta, y, z if (spec.relu and res is not None) else None, mean, rstd)
        # fused BatchNorm-backward reduction: `producer` = link of the layer that made x (this conv is its only
        # consumer); `link` = what THIS layer offers to its own single consumer (no residual, per-rank statistics)
        ctx.producer = producer
        ctx.link = None
        ConvBNAct.last_link = None
        if FUSE_BNRED and res is None and world == 1:
            # (`z is y` when the apply pass is deferred: the link then holds a detached alias -- the returned tensor carries
            # the link as an attribute, and a link that pointed back at it would be a reference cycle that keeps this step's
            # autograd graph, with the streams its AccumulateGrad nodes were created on, alive into the next capture)
            ctx.link = ConvBNAct.last_link = _BnLink(y.detach() if z is y else y, mean, rstd, gamma, beta, spec.relu, scratch, bn, C)
        return z

    @staticmethod
    def backward(ctx, dz):
        spec = ctx.spec
        x, weight, gamma, beta, y, z, mean, rstd = ctx.saved_tensors
        dz = ops.as_nhwc(dz)
        C = weight.shape[0]
        gw, gg, gb = ctx.arena       # FlatAdamW gradient-arena views: accumulate in place
        if gg is None or gb is None:
            dgb = ops.zeros_f32(2 * C, weight.device)
            gg_out, gb_out = dgb[:C], dgb[C:]
        else:
            gg_out, gb_out = gg, gb
        want_dres = ctx.has_res and ctx.needs_input_grad[1]
        link = ctx.link
        prod = ctx.producer if (ctx.producer is not None and ctx.producer.usable()) else None
        dw = gw if gw is not None else torch.zeros_like(weight)
        wpT = ctx.packed[1] if (spec.kind == 'pw' and ctx.packed is not None) else None
        if (FUSE_BNAPPLY and spec.kind == 'pw' and wpT is not None and spec.impl == 1 and not ctx.has_res
                and ctx.sync[0] == 1 and ctx.needs_input_grad[0] and dz.dtype == torch.bfloat16
                and C % 8 == 0 and weight.shape[1] % 16 == 0):
            # one kernel: this layer's BatchNorm-backward apply -> dgrad (-> the producer's reduction)
            sums, mask = _finished_sums(ctx, dz)
            dy, dx = ops.pwconv_bwd_fused(dz, y, mean, rstd, gamma, beta, sums, mask, wpT, dgamma=gg_out, dbeta=gb_out,
                                          link=prod)
            if prod is not None:
                prod.reduced, prod.bn._tss_dirty = True, True
            if gw is not None:
                wgrad_lane.run(dy.device, lambda: ops.pwconv_wgrad(x, dy, dw, impl=1), x, dy)
            else:
                ops.pwconv_wgrad(x, dy, dw, impl=1)
            return _conv_grads(ctx, dx, None, dw, gg_out, gb_out)
        if (STEM_TC and STEM_BWD_FUSED and spec.kind == 'stem' and dz.dtype == torch.bfloat16 and C == 32 and not ctx.has_res
                and ctx.sync[0] == 1 and not ctx.needs_input_grad[0] and ops.geom(dz)[4] == C):
            # the stem has no input gradient: BatchNorm-backward apply + weight gradient in one kernel, no dy tensor
            sums, mask = _finished_sums(ctx, dz)
            fused = lambda: ops.stem_wgrad_tc_bn(x, dz, y, mean, rstd, gamma, beta, sums, mask, dw, dgamma=gg_out, dbeta=gb_out)
            if gw is not None:
                wgrad_lane.run(dz.device, fused, x, dz, y, sums)
            else:
                fused()
            return _conv_grads(ctx, None, None, dw, gg_out, gb_out)
        if (FUSE_BNAPPLY_DW and spec.kind == 'dw' and spec.stride == 1 and spec.dilation == 1
                and C % 32 == 0 and not ctx.has_res and ctx.sync[0] == 1 and ctx.needs_input_grad[0]
                and ops.geom(dz)[4] == C):
            # one kernel: this layer's BatchNorm-backward apply -> depthwise dgrad -> the producer's reduction
            sums, mask = _finished_sums(ctx, dz)
            dy, dx = ops.dwconv_bwd_fused(dz, y, weight, mean, rstd, gamma, beta, sums, mask, prod, dgamma=gg_out, dbeta=gb_out)
            if prod is not None:
                prod.reduced, prod.bn._tss_dirty = True, True
            dw_wgrad = _dw_wgrad_fn(ctx, x, dy, dw, spec)
            if gw is not None:
                wgrad_lane.run(dy.device, dw_wgrad, x, dy)
            else:
                dw_wgrad()
            return _conv_grads(ctx, dx, None, dw, gg_out, gb_out)
        if link is not None and link.reduced:
            # the consumer's dgrad already masked the gradient and accumulated both sums
            dy, dres = ops.bn_backward(dz, None, y, mean, rstd, gamma, False, dgamma=gg_out, dbeta=gb_out, beta=beta,
                                       sums=link.sums, prereduced=True)
        else:
            sums = ctx.scratch[2 * C:].view(torch.float32)
            if spec.bn._tss_dirty or ctx.scratch is not getattr(spec.bn, '_tss_scratch', None):
                sums = None      # a second backward without a forward in between: fresh zeros
            spec.bn._tss_dirty = True
            dy, dres = ops.bn_backward(dz, z, y, mean, rstd, gamma, spec.relu, want_dres=want_dres,
                                       dgamma=gg_out, dbeta=gb_out, beta=beta, sums=sums, sync=ctx.sync)
        dx = None
        # wgrad off the critical chain when it accumulates in place into the optimizer's arena
        lane = (lambda fn: wgrad_lane.run(dy.device, fn, x, dy)) if gw is not None else (lambda fn: fn())
        if spec.kind == 'pw':
            impl = spec.impl if wpT is not None else 0
            lane(lambda: ops.pwconv_wgrad(x, dy, dw, impl=impl))
            if ctx.needs_input_grad[0]:
                if prod is not None and impl == 1 and weight.shape[1] % 16 == 0:
                    dx = ops.pwconv_dgrad_bnred(dy, wpT, prod)
                    prod.reduced, prod.bn._tss_dirty = True, True
                else:
                    dx = ops.pwconv_dgrad(dy, weight, wpT=wpT, impl=impl)
        elif spec.kind == 'dw':
            lane(_dw_wgrad_fn(ctx, x, dy, dw, spec))
            if ctx.needs_input_grad[0]:
                if prod is not None and spec.stride == 1 and spec.dilation == 1 and weight.shape[0] % 32 == 0:
                    dx = ops.dwconv_dgrad_bnred(dy, weight, prod)
                    prod.reduced, prod.bn._tss_dirty = True, True
                elif prod is not None and FUSE_BNRED_EXT and spec.stride == 2 and spec.dilation == 1:
                    dx = ops.dwconv_dgrad_s2_bnred(dy, weight, prod)
                    prod.reduced, prod.bn._tss_dirty = True, True
                else:
                    dx = ops.dwconv_dgrad(dy, weight, ctx.in_hw[0], ctx.in_hw[1], spec.stride, spec.dilation)
        else:  # stem: the image needs no gradient
            if ctx.needs_input_grad[0]:
                raise RuntimeError('gradient w.r.t. the input image is not implemented')
            if ctx.patches is not None and dy.dtype == torch.bfloat16:
                patches = ctx.patches        # made on the lane's stream in forward: stream order is the dependency
                lane(lambda: ops.stem_wgrad_from_patches(patches, dy, dw))
            elif STEM_TC and dy.dtype == torch.bfloat16 and weight.shape[0] == 32:
                lane(lambda: ops.stem_wgrad_tc(x, dy, dw))
            else:
                lane(lambda: ops.stem_wgrad(x, dy, dw))
        return _conv_grads(ctx, dx, dres, dw, gg_out, gb_out)


class Im2Col3x3(torch.autograd.Function):
    """Patch matrix of a dense 3x3 / stride 1 / padding 1 convolution (contextnet.py:55)."""

    @staticmethod
    def forward(ctx, x):
        return ops.im2col3x3(x)

    @staticmethod
    def backward(ctx, dcol):
        return ops.col2im3x3(ops.as_nhwc(dcol))


class TapMajorWeight(torch.autograd.Function):
    """(Cout, Cin, 3, 3) parameter -> the (Cout, 9*Cin, 1, 1) matrix the patch GEMM multiplies with;
    backward accumulates straight into the optimizer's gradient arena when there is one."""

    @staticmethod
    def forward(ctx, w):
        ctx.param = w
        ctx.arena = getattr(w, '_tss_grad', None)
        ctx.shape = w.shape
        return ops.permute_weights3x3(w)

    @staticmethod
    def backward(ctx, dwk):
        if ctx.arena is not None:
            ops.permute_weights3x3_bwd(dwk.contiguous(), ctx.arena)
            grad_ready(ctx.param)
            return None
        dw = torch.zeros(ctx.shape, dtype=torch.float32, device=dwk.device)
        return ops.permute_weights3x3_bwd(dwk.contiguous(), dw)


# The class-score conv (128 -> 19 with bias) on the tcgen05 GEMMs with zero-padded operands instead of the SIMT GEMMs
# (forward 47 us, dgrad 29 us, wgrad 62 us at 12 x 96 x 96 x 128 on B200: the three slowest launches per byte of the step).
CLASS_TC = gate('CLASS_TC')


class ConvBias(torch.autograd.Function):
    """y = conv1x1(x, w) + b  (class scores; Nc = 19 lives in a padded channel pitch)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        Nc, K = weight.shape[0], weight.shape[1]
        ctx.packed = None
        if CLASS_TC and x.dtype == torch.bfloat16 and x.is_cuda and K % 16 == 0 and Nc % 16 != 0 and Nc <= 32 \
                and ops.geom(x) is not None:
            ctx.packed = ops.class_scores_pack(weight.detach(), bias.detach() if bias is not None else None)
            y = ops.class_scores_fwd(x, weight, ctx.packed)
        else:
            y = ops.pwconv_fwd(x, weight, shift=bias)
        ctx.save_for_backward(x, weight, bias)
        ctx.params = (weight, bias)
        ctx.arena = (getattr(weight, '_tss_grad', None),
                     getattr(bias, '_tss_grad', None) if bias is not None else None)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        gw, gbias = ctx.arena
        Nc = weight.shape[0]
        g = ops.geom(dy)
        if g is None or g[4] % 8 != 0 or g[4] < (Nc + 7) // 8 * 8:
            # re-pitch: gradient rows must be 16-byte aligned with zeroed pad columns
            pitch = (Nc + 7) // 8 * 8 + 8
            buf = torch.zeros((dy.shape[0], dy.shape[2], dy.shape[3], pitch), dtype=dy.dtype, device=dy.device)
            view = buf[..., :Nc].permute(0, 3, 1, 2)
            view.copy_(dy)
            dy = view
        dw = gw if gw is not None else torch.zeros_like(weight)
        db = None
        if ctx.has_bias:
            db = gbias if gbias is not None else torch.zeros(Nc, dtype=torch.float32, device=weight.device)
        if ctx.packed is not None and dy.dtype == torch.bfloat16:
            # pad columns of dy (Nc .. pitch) are zeros by construction (fused head / the re-pitch above)
            dx = ops.class_scores_dgrad(dy, weight, ctx.packed) if ctx.needs_input_grad[0] else None
            ops.pwconv_wgrad(x, dy, dw, db, impl=1)
        else:
            dx = ops.pwconv_dgrad(dy, weight) if ctx.needs_input_grad[0] else None
            ops.pwconv_wgrad(x, dy, dw, db)
        grad_ready(*ctx.params)
        return dx, None if gw is not None else dw, None if gbias is not None else db


class AdaptivePool(torch.autograd.Function):
    """All pyramid bins in one pass; outputs one (N,C,b,b) tensor per bin."""

    @staticmethod
    def forward(ctx, x, bins):
        ctx.bins, ctx.shape = tuple(bins), x.shape
        ctx.dtype, ctx.device = x.dtype, x.device
        _, outs = ops.adaptive_pool_fwd(x, ctx.bins)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        N, C, H, W = ctx.shape
        dev = ctx.device
        cells = sum(b * b for b in ctx.bins)
        dbuf = torch.zeros((cells * N, C), dtype=ctx.dtype, device=dev)
        for view, g in zip(ops.split_pool_buffer(dbuf, N, C, ctx.bins), grads):
            if g is not None:
                ops.copy_rows(ops.as_nhwc(g), view)
        dx = ops.empty_nhwc(N, C, H, W, ctx.dtype, dev)
        ops.adaptive_pool_bwd(dbuf, dx, ctx.bins, accumulate=False)
        return dx, None


class PPMConcat(torch.autograd.Function):
    """cat([x, up(z_1), ..., up(z_k)], dim=1) with the bilinear up-samplings written
    straight into their channel slices of the concat buffer (fastscnn.py:119-122)."""

    @staticmethod
    def forward(ctx, x, *zs):
        N, C, H, W = x.shape
        total = C + sum(z.shape[1] for z in zs)
        cat = ops.empty_nhwc(N, total, H, W, x.dtype, x.device)
        ops.copy_rows(x, cat[:, :C])
        off = C
        for z in zs:
            ops.bilinear_fwd(z, H, W, out=cat[:, off:off + z.shape[1]])
            off += z.shape[1]
        ctx.c_in = C
        ctx.z_shapes = [z.shape for z in zs]
        return cat

    @staticmethod
    def backward(ctx, dcat):
        dcat = ops.as_nhwc(dcat)
        C = ctx.c_in
        N, _, H, W = dcat.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.copy_rows(dcat[:, :C], ops.empty_nhwc(N, C, H, W, dcat.dtype, dcat.device))
        dzs, off = [], C
        for shp in ctx.z_shapes:
            dzs.append(ops.bilinear_bwd(dcat[:, off:off + shp[1]], shp[2], shp[3]))
            off += shp[1]
        return (dx, *dzs)


class PPMBranches(torch.autograd.Function):
    """cat([x, up(relu(bn(conv1x1(pool_b(x)))))...]) for all pyramid bins (fastscnn.py:101-122) with grouped
    kernels.  ``state`` carries the bins, the BatchNorm hyper-parameters and the device table of parameter /
    buffer / gradient-arena addresses; ``params`` (weight, gamma, beta per branch) are only listed so that
    autograd knows the node and the reducer hears about their gradients."""

    @staticmethod
    def forward(ctx, x, state, *params):
        N, C, H, W = x.shape
        Cb = params[0].shape[0]
        pool, _ = ops.adaptive_pool_fwd(x, state.bins)
        y, z, mean, rstd = ops.ppm_branches_fwd(pool, state.table, N, C, Cb, state.bins, state.momentum, state.eps)
        cat = ops.ppm_concat_fwd(x, z, Cb, state.bins)
        ctx.save_for_backward(pool, y, mean, rstd)
        ctx.state, ctx.params, ctx.shape = state, params, (N, C, H, W, Cb)
        return cat

    @staticmethod
    def backward(ctx, dcat):
        pool, y, mean, rstd = ctx.saved_tensors
        N, C, H, W, Cb = ctx.shape
        st = ctx.state
        dcat = ops.as_nhwc(dcat)
        dz = ops.ppm_concat_bwd(dcat, C, Cb, st.bins)
        dpool = ops.ppm_branches_bwd(dz, y, pool, st.table, mean, rstd, N, st.bins)
        dx = ops.copy_rows(dcat[:, :C], ops.empty_nhwc(N, C, H, W, dcat.dtype, dcat.device))
        ops.adaptive_pool_bwd(dpool, dx, st.bins, accumulate=True)
        grad_ready(*ctx.params)
        return (dx, None) + (None,) * len(ctx.params)


# nn.Dropout of the classifiers on the library's own kernel (csrc/dropout.cu: counter-based mask, regenerated in
# the backward pass, no mask tensor, no ATen kernels on the path).  On by default (B200: 28.8 -> 14.1 us per pass).
OWN_DROPOUT = gate('OWN_DROPOUT')


class Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        y, used = ops.dropout_fwd(ops.as_nhwc(x), p)
        ctx.p = p
        ctx.save_for_backward(used)
        return y

    @staticmethod
    def backward(ctx, dy):
        (used,) = ctx.saved_tensors
        return ops.dropout_bwd(ops.as_nhwc(dy).contiguous(memory_format=torch.channels_last), ctx.p, used), None


class Bilinear:
    """``F.interpolate(mode='bilinear', align_corners=True)`` on NHWC tensors: the registered, differentiable operator
    ``torch.ops.tss_b200.bilinear`` (library.py: schema, fake implementation, autograd formula)."""

    @staticmethod
    def apply(x, Ho, Wo):
        return torch.ops.tss_b200.bilinear(x, Ho, Wo)


class UpsampleLogits:
    """NHWC class scores -> NCHW-contiguous full-resolution logits (the reference's output layout):
    ``torch.ops.tss_b200.upsample_logits``."""

    @staticmethod
    def apply(x, Ho, Wo):
        return torch.ops.tss_b200.upsample_logits(x, Ho, Wo)


class UpsampleCrossEntropy(torch.autograd.Function):
    """mean CE of the x8-upsampled class scores, computed from the 1/8-resolution scores: the logits
    are interpolated on the fly and the gradient is folded through the interpolation's transpose in
    the same pass (fastscnn.py:63-64 + train_fastscnn.py:132 as ONE kernel)."""

    @staticmethod
    def forward(ctx, scores, target, ignore_index, Ho, Wo):
        want = ctx.needs_input_grad[0]
        loss, dx, _ = ops.upsample_ce_forward(scores, target, Ho, Wo, ignore_index, want_grad=want)
        ctx.save_for_backward(dx)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dx,) = ctx.saved_tensors
        if dx is None:
            return None, None, None, None, None
        if not _unit_grad(ctx):
            dx = dx * grad_out.to(dx.dtype)       # 1/64 of the logits' size: not worth a kernel
        return dx, None, None, None, None


class PixelCrossEntropy(torch.autograd.Function):
    """``F.cross_entropy(..., reduction='none')`` (losses/ohem_loss.py:11-12): per-pixel loss, 0 at
    ignored pixels.  Backward recomputes the softmax (one more pass over the logits)."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        pixel = ops.ce_forward(logits, target, ignore_index, want_grad=False, want_pixel_loss=True)[2]
        ctx.save_for_backward(logits, target)
        ctx.ignore_index = ignore_index
        return pixel

    @staticmethod
    def backward(ctx, gpix):
        logits, target = ctx.saved_tensors
        # d loss_i / d logits_i = softmax - onehot; computed with unit weight (nvalid = 1), then scaled per pixel
        one = torch.ones(1, dtype=torch.int64, device=logits.device)
        N, C, H, W = logits.shape
        dl = torch.empty_like(logits, memory_format=torch.contiguous_format)
        from . import _lib
        _lib.call('tss_ce_fwd', logits=logits.contiguous(), target=target.long().contiguous(), N=N, C=C, HW=H * W,
                  ignore_index=ctx.ignore_index, nvalid=one, loss_sum=None, pixel_loss=None, dlogits=dl, ohem=None,
                  dtype=_lib.dtype_code(logits.dtype))
        return dl * gpix.unsqueeze(1).to(dl.dtype), None, None


class OhemCrossEntropy(torch.autograd.Function):
    """``ohem_loss`` (losses/ohem_loss.py:10-21) on device: per-pixel CE -> radix select of the
    (n+1)-th largest loss -> case decision -> OHEM-weighted gradient, with no sort and no host
    synchronisation.  ``source`` is either the NCHW logits or, for the untouched output of one of this
    package's models, the 1/8-resolution scores they were interpolated from (fused head)."""

    @staticmethod
    def forward(ctx, source, target, ignore_index, thresh, numel_frac, out_hw):
        fused = out_hw is not None
        if fused:
            Ho, Wo = out_hw
            pixel = ops.upsample_ce_forward(source, target, Ho, Wo, ignore_index, want_grad=False, want_pixel_loss=True)[2]
        else:
            pixel = ops.ce_forward(source, target, ignore_index, want_grad=False, want_pixel_loss=True)[2]
        n = pixel.numel()
        loss, rule = ops.ohem_select(pixel, int(n * numel_frac), thresh)
        grad = None
        if ctx.needs_input_grad[0]:
            if fused:
                grad = ops.upsample_ce_forward(source, target, Ho, Wo, ignore_index, want_grad=True, ohem=rule)[1]
            else:
                grad = ops.ce_forward(source, target, ignore_index, want_grad=True, ohem=rule)[1]
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None, None, None, None, None
        if not _unit_grad(ctx):
            grad = grad * grad_out.to(grad.dtype)
        return grad, None, None, None, None, None


def attach_head(logits, scores):
    """Remember on the model's output which low-resolution scores it was interpolated from, so that
    ``losses.cross_entropy`` can take the fused head instead of re-reading the logits."""
    logits._tss_head = (scores, logits._version)
    return logits


# Training with this package's losses: the full-resolution logits (269 MB of bf16 per 12 x 768 x 768 step) are only
# ever consumed by the fused head, which works from the 1/8 class scores.  With defer_logits set on the model (the
# trainer does it when the loss declares ``accepts_deferred_logits``; gate DEFER_LOGITS, on by default) the training
# forward returns this handle instead of launching the x8 up-sampling;
# anything else that wants the tensor calls ``materialize()``.
DEFER_LOGITS = gate('DEFER_LOGITS')


class DeferredLogits:
    """(N, C, 8h, 8w) logits that have not been interpolated from their (N, C, h, w) NHWC ``scores`` yet."""

    __slots__ = ('scores', 'height', 'width')

    def __init__(self, scores, height, width):
        self.scores, self.height, self.width = scores, height, width

    @property
    def shape(self):
        return torch.Size((self.scores.shape[0], self.scores.shape[1], self.height, self.width))

    def dim(self):
        return 4

    def materialize(self):
        return attach_head(UpsampleLogits.apply(self.scores, self.height, self.width), self.scores)


def model_output(module, scores, height, width):
    """What a model's forward returns for the class scores: the up-sampled logits, or the deferred handle in
    training mode when the module opted in."""
    if module.training and getattr(module, 'defer_logits', False) and torch.is_grad_enabled():
        return DeferredLogits(scores, height, width)
    return attach_head(UpsampleLogits.apply(scores, height, width), scores)


def enable_deferred_logits(model, loss_fn):
    """Trainer / bench hook: switch the model to deferred logits iff the gate is on and the loss can take them."""
    on = bool(DEFER_LOGITS and getattr(loss_fn, 'accepts_deferred_logits', False) and hasattr(model, 'defer_logits'))
    if hasattr(model, 'defer_logits'):
        model.defer_logits = on
    return on


def fused_head_source(logits):
    """-> the NHWC scores ``logits`` was up-sampled from, if it is still the untouched model output."""
    if isinstance(logits, DeferredLogits):
        return logits.scores
    head = getattr(logits, '_tss_head', None)
    if head is None or head[1] != logits._version:
        return None
    return head[0]


class CrossEntropy(torch.autograd.Function):
    """mean softmax cross-entropy over non-ignored pixels; gradient computed in the forward pass."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        want = ctx.needs_input_grad[0]
        loss, dlogits, _, _ = ops.ce_forward(logits, target, ignore_index, want_grad=want)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dlogits,) = ctx.saved_tensors
        if dlogits is None:
            return None, None, None
        if not _unit_grad(ctx):
            # out of place: the saved gradient stays the unit-weight one if backward runs again (retain_graph)
            dlogits = dlogits * grad_out.to(dlogits.dtype)
        return dlogits, None, None
