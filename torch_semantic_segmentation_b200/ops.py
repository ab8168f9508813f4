"""Tensor-level wrappers of the C-ABI kernels (``include/tss_b200.h``).

Activations are 4-D tensors with *logical* shape (N, C, H, W) and *physical* dense NHWC
layout (``torch.channels_last``), optionally with a channel pitch ``ld >= C`` (a channel
slice of a wider NHWC buffer).  Wrappers allocate outputs through PyTorch's caching
allocator, derive sizes/pitches from the tensors and call the library on the current
stream.  Nothing here computes on the host and nothing falls back to torch ops.
"""
import os

import torch

from . import _lib
from .gates import gate
from ._lib import EPI_RELU, dtype_code

PYRAMID_BINS = (1, 2, 3, 6)

# Bumped whenever a kernel rewrites parameters or BatchNorm buffers behind autograd's back
# (raw-pointer writes do not touch tensor._version); derived caches key on it.
WEIGHTS_EPOCH = [0]


# ------------------------------------------------------------------ layout helpers -----
def empty_nhwc(N, C, H, W, dtype, device, pitch=None):
    """Logical (N,C,H,W) tensor over a dense NHWC buffer (channel pitch ``pitch``)."""
    ld = C if pitch is None else pitch
    buf = torch.empty((N, H, W, ld), dtype=dtype, device=device)
    return buf[..., :C].permute(0, 3, 1, 2)


def geom(t):
    """-> (N, C, H, W, ld) of an NHWC(-pitched) activation, or None if it is not one."""
    if t.dim() != 4:
        return None
    N, C, H, W = t.shape
    sN, sC, sH, sW = t.stride()
    if W > 1:
        ld = sW
    elif H > 1:
        ld = sH
    elif N > 1:
        ld = sN
    else:
        ld = C
    ok = (C == 1 or sC == 1) and ld >= C and (W == 1 or sW == ld) and \
        (H == 1 or sH == W * ld) and (N == 1 or sN == H * W * ld)
    return (N, C, H, W, ld) if ok else None


def as_nhwc(t):
    """Return ``t`` if it already is NHWC(-pitched), else a dense NHWC copy (API boundary only)."""
    if geom(t) is not None:
        return t
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def _g(t, name):
    g = geom(t)
    if g is None:
        raise RuntimeError('%s: tensor of shape %s / strides %s is not NHWC' % (name, tuple(t.shape), t.stride()))
    return g


def _flags(relu):
    return EPI_RELU if relu else 0


def zeros_f32(n, device):
    return torch.zeros(n, dtype=torch.float32, device=device)


def zeros_f64(n, device):
    return torch.zeros(n, dtype=torch.float64, device=device)


# ------------------------------------------------------------------ depthwise 3x3 ------
def dwconv_fwd(x, w, stride, dilation, scale=None, shift=None, relu=False, stats=None):
    N, C, Hi, Wi, ld = _g(x, 'dwconv_fwd')
    if ld != C:
        raise RuntimeError('dwconv_fwd: pitched input not supported')
    Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
    y = empty_nhwc(N, C, Ho, Wo, x.dtype, x.device)
    _lib.call('tss_dwconv3x3_fwd', x=x, w=w, y=y, N=N, Hi=Hi, Wi=Wi, C=C, stride=stride,
              dilation=dilation, scale=scale, shift=shift, flags=_flags(relu), stats=stats,
              dtype=dtype_code(x.dtype))
    return y


def dwconv_dgrad(dy, w, Hi, Wi, stride, dilation):
    N, C, Ho, Wo, ld = _g(dy, 'dwconv_dgrad')
    if ld != C:
        raise RuntimeError('dwconv_dgrad: pitched input not supported')
    dx = empty_nhwc(N, C, Hi, Wi, dy.dtype, dy.device)
    _lib.call('tss_dwconv3x3_dgrad', dy=dy, w=w, dx=dx, N=N, Hi=Hi, Wi=Wi, C=C, stride=stride,
              dilation=dilation, dtype=dtype_code(dy.dtype))
    return dx


def dwconv_wgrad(x, dy, dw, stride, dilation):
    """dw (fp32, shape (C,1,3,3)) += wgrad."""
    N, C, Hi, Wi, ld = _g(x, 'dwconv_wgrad')
    if ld != C or _g(dy, 'dwconv_wgrad')[4] != C:
        raise RuntimeError('dwconv_wgrad: pitched input not supported')
    _lib.call('tss_dwconv3x3_wgrad', x=x, dy=dy, dw=dw, N=N, Hi=Hi, Wi=Wi, C=C, stride=stride,
              dilation=dilation, dtype=dtype_code(x.dtype))


def dwconv_fwd_bnin(x, in_scale, in_shift, in_relu, w, stride, stats):
    """Depthwise forward on ``act(x * in_scale + in_shift)`` without materialising it (``x``: raw conv output)."""
    N, C, Hi, Wi, ld = _g(x, 'dwconv_fwd_bnin')
    if ld != C:
        raise RuntimeError('dwconv_fwd_bnin: pitched input not supported')
    Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
    y = empty_nhwc(N, C, Ho, Wo, x.dtype, x.device)
    _lib.call('tss_dwconv3x3_fwd_bnin', x=x, in_scale=in_scale, in_shift=in_shift, in_flags=_flags(in_relu), w=w, y=y, N=N,
              Hi=Hi, Wi=Wi, C=C, stride=stride, stats=stats, dtype=dtype_code(x.dtype))
    return y


def dwconv_wgrad_bnin(x, in_scale, in_shift, in_relu, dy, dw, stride):
    N, C, Hi, Wi, ld = _g(x, 'dwconv_wgrad_bnin')
    if ld != C or _g(dy, 'dwconv_wgrad_bnin')[4] != C:
        raise RuntimeError('dwconv_wgrad_bnin: pitched input not supported')
    _lib.call('tss_dwconv3x3_wgrad_bnin', x=x, in_scale=in_scale, in_shift=in_shift, in_flags=_flags(in_relu), dy=dy, dw=dw,
              N=N, Hi=Hi, Wi=Wi, C=C, stride=stride, dtype=dtype_code(x.dtype))


# ------------------------------------------------------------------ pointwise 1x1 ------
def _pad8(c):
    return (c + 7) // 8 * 8


def pwconv_fwd(x, w, scale=None, shift=None, res=None, relu=False, stats=None, out=None,
               wp=None, impl=0):
    """y = x . w^T (+epilogue).  ``w`` fp32 (Nc, K, 1, 1).  Nc % 8 != 0 gets a padded pitch."""
    N, K, H, W, ldx = _g(x, 'pwconv_fwd')
    Nc = w.shape[0]
    if out is None:
        out = empty_nhwc(N, Nc, H, W, x.dtype, x.device, pitch=None if Nc % 8 == 0 else _pad8(Nc) + 8)
    ldy = _g(out, 'pwconv_fwd')[4]
    ldr = _g(res, 'pwconv_fwd')[4] if res is not None else 0
    _lib.call('tss_pwconv_fwd', x=x, w=w, wp=wp, y=out, M=N * H * W, K=K, Nc=Nc, ldx=ldx, ldy=ldy,
              scale=scale, shift=shift, res=res, ldr=ldr, flags=_flags(relu), stats=stats,
              impl=impl, dtype=dtype_code(x.dtype))
    return out


def pwconv_fwd_bnin(x, in_scale, in_shift, in_relu, wp, stats):
    """1x1 forward on ``act(x * in_scale + in_shift)`` built inside the GEMM (``x``: raw conv output, bf16)
    -> y (raw output), z (the activated input, stored once for the weight gradient)."""
    N, K, H, W, ldx = _g(x, 'pwconv_fwd_bnin')
    Nc = wp.shape[0]
    y = empty_nhwc(N, Nc, H, W, x.dtype, x.device)
    z = empty_nhwc(N, K, H, W, x.dtype, x.device)
    _lib.call('tss_pwconv_fwd_bnin', x=x, ldx=ldx, in_scale=in_scale, in_shift=in_shift, in_flags=_flags(in_relu), z=z, ldz=K,
              wp=wp, y=y, ldy=Nc, M=N * H * W, K=K, Nc=Nc, stats=stats)
    return y, z


def pwconv_dgrad(dy, w, wpT=None, impl=0):
    N, Nc, H, W, lddy = _g(dy, 'pwconv_dgrad')
    K = w.shape[1]
    dx = empty_nhwc(N, K, H, W, dy.dtype, dy.device)
    _lib.call('tss_pwconv_dgrad', dy=dy, w=w, wpT=wpT, dx=dx, M=N * H * W, K=K, Nc=Nc, lddy=lddy,
              lddx=K, impl=impl, dtype=dtype_code(dy.dtype))
    return dx


def pwconv_dgrad_bnred(dy, wpT, link):
    """tcgen05 dgrad + the producer's BatchNorm-backward reduction (``link``: functional._BnLink) -> g."""
    N, Nc, H, W, lddy = _g(dy, 'pwconv_dgrad_bnred')
    K = wpT.shape[0]
    g = empty_nhwc(N, K, H, W, dy.dtype, dy.device)
    _lib.call('tss_pwconv_dgrad_bnred', dy=dy, wpT=wpT, g=g, M=N * H * W, K=K, Nc=Nc, lddy=lddy, ldg=K, yp=link.y,
              ldyp=_g(link.y, 'pwconv_dgrad_bnred')[4], mean=link.mean, rstd=link.rstd, gamma=link.gamma,
              beta=link.beta, flags=_flags(link.relu), sums=link.sums)
    return g


def pwconv_bwd_fused(dz, y, mean, rstd, gamma, beta, sums, relu, wpT, dgamma=None, dbeta=None, link=None):
    """BatchNorm-backward apply + tcgen05 dgrad (+ the producer's reduction, ``link``) in one kernel -> dy, dx.
    ``sums`` is the finished reduction of this layer; ``relu``: the mask is recomputed from ``y``."""
    N, Nc, H, W, lddz = _g(dz, 'pwconv_bwd_fused')
    K = wpT.shape[0]
    M = N * H * W
    dy = empty_nhwc(N, Nc, H, W, dz.dtype, dz.device)
    dx = empty_nhwc(N, K, H, W, dz.dtype, dz.device)
    _lib.call('tss_pwconv_bwd_fused', dz=dz, y=y, lddz=lddz, ldy=_g(y, 'pwconv_bwd_fused')[4], mean=mean, rstd=rstd,
              gamma=gamma, beta=beta, sums=sums, flags=_flags(relu), count=M, dy=dy, lddy=Nc, dgamma=dgamma,
              dbeta=dbeta, wpT=wpT, dx=dx, M=M, K=K, Nc=Nc, lddx=K,
              yp=link.y if link is not None else None,
              ldyp=_g(link.y, 'pwconv_bwd_fused')[4] if link is not None else 0,
              pmean=link.mean if link is not None else None, prstd=link.rstd if link is not None else None,
              pgamma=link.gamma if link is not None else None, pbeta=link.beta if link is not None else None,
              pflags=_flags(link.relu) if link is not None else 0, psums=link.sums if link is not None else None)
    return dy, dx


def bn_backward_reduce(dz, y, mean, rstd, gamma, beta, relu, sums):
    """The reduction pass of ``bn_backward`` alone (mask recomputed from ``y``); ``sums`` zeroed by the caller."""
    N, C, H, W, lddz = _g(dz, 'bn_backward_reduce')
    _lib.call('tss_bn_bwd_reduce', dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums,
              M=N * H * W, C=C, lddz=lddz, ldz=0, ldy=_g(y, 'bn_backward_reduce')[4], flags=_flags(relu),
              dtype=dtype_code(dz.dtype))


def dwconv_dgrad_bnred(dy, w, link):
    """stride-1 depthwise dgrad + the producer's BatchNorm-backward reduction -> g."""
    N, C, H, W, ld = _g(dy, 'dwconv_dgrad_bnred')
    if ld != C or _g(link.y, 'dwconv_dgrad_bnred')[4] != C:
        raise RuntimeError('dwconv_dgrad_bnred: pitched tensors not supported')
    g = empty_nhwc(N, C, H, W, dy.dtype, dy.device)
    _lib.call('tss_dwconv3x3_dgrad_bnred', dy=dy, w=w, g=g, N=N, H=H, W=W, C=C, yp=link.y, mean=link.mean,
              rstd=link.rstd, gamma=link.gamma, beta=link.beta, flags=_flags(link.relu), sums=link.sums,
              dtype=dtype_code(dy.dtype))
    return g


def dwconv_bwd_fused(dz, y, w, mean, rstd, gamma, beta, sums, relu, link, dgamma=None, dbeta=None):
    """BatchNorm-backward apply + stride-1 depthwise dgrad (+ the producer's reduction when there is a ``link``)
    -> dy, g."""
    N, C, H, W, ld = _g(dz, 'dwconv_bwd_fused')
    if ld != C or _g(y, 'dwconv_bwd_fused')[4] != C or (link is not None and _g(link.y, 'dwconv_bwd_fused')[4] != C):
        raise RuntimeError('dwconv_bwd_fused: pitched tensors not supported')
    dy = empty_nhwc(N, C, H, W, dz.dtype, dz.device)
    g = empty_nhwc(N, C, H, W, dz.dtype, dz.device)
    lk = link
    _lib.call('tss_dwconv3x3_bwd_fused', dz=dz, y=y, w=w, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums,
              flags=_flags(relu), count=N * H * W, dy=dy, dgamma=dgamma, dbeta=dbeta, g=g, N=N, H=H, W=W, C=C,
              yp=lk.y if lk else None, pmean=lk.mean if lk else None, prstd=lk.rstd if lk else None,
              pgamma=lk.gamma if lk else None, pbeta=lk.beta if lk else None, pflags=_flags(lk.relu) if lk else 0,
              psums=lk.sums if lk else None, dtype=dtype_code(dz.dtype))
    return dy, g


def dwconv_dgrad_s2_bnred(dy, w, link):
    """stride-2 depthwise dgrad + the producer's BatchNorm-backward reduction -> g (the producer's geometry)."""
    N, C, Hi, Wi, ld = _g(link.y, 'dwconv_dgrad_s2_bnred')
    if ld != C or _g(dy, 'dwconv_dgrad_s2_bnred')[4] != C:
        raise RuntimeError('dwconv_dgrad_s2_bnred: pitched tensors not supported')
    g = empty_nhwc(N, C, Hi, Wi, dy.dtype, dy.device)
    _lib.call('tss_dwconv3x3_dgrad_s2_bnred', dy=dy, w=w, g=g, N=N, Hi=Hi, Wi=Wi, C=C, yp=link.y, mean=link.mean,
              rstd=link.rstd, gamma=link.gamma, beta=link.beta, flags=_flags(link.relu), sums=link.sums,
              dtype=dtype_code(dy.dtype))
    return g


def pwconv_wgrad(x, dy, dw, db=None, impl=0):
    """dw (fp32 (Nc,K,1,1)) += dy^T x ; db (fp32 (Nc,)) += colsum(dy)."""
    N, K, H, W, ldx = _g(x, 'pwconv_wgrad')
    Nc, lddy = dy.shape[1], _g(dy, 'pwconv_wgrad')[4]
    _lib.call('tss_pwconv_wgrad', x=x, dy=dy, dw=dw, db=db, M=N * H * W, K=K, Nc=Nc, ldx=ldx,
              lddy=lddy, impl=impl, dtype=dtype_code(x.dtype))


def class_scores_pack(w, bias, grad_pitch=None):
    """Zero-padded tensor-core operands of the class-score conv (Nc not a multiple of 16, with bias):
    -> wp (Np, K) bf16, wpT (K, Npt) bf16, bias_pad (Np,) fp32; Npt = channel pitch of the gradient that will come back."""
    Nc, K = w.shape[0], w.shape[1]
    Np = (Nc + 15) // 16 * 16
    Npt = _pad8(Nc) if grad_pitch is None else int(grad_pitch)
    wp = torch.empty((Np, K), dtype=torch.bfloat16, device=w.device)
    wpT = torch.empty((K, Npt), dtype=torch.bfloat16, device=w.device)
    bpad = torch.empty(Np, dtype=torch.float32, device=w.device)
    _lib.call('tss_class_scores_pack', w=w, bias=bias, wp=wp, wpT=wpT, bias_pad=bpad, Nc=Nc, K=K, Np=Np, Npt=Npt)
    return wp, wpT, bpad


def class_scores_fwd(x, w, packed):
    """(N, Nc, H, W) scores in a padded channel pitch, on the tcgen05 GEMM (``packed`` from class_scores_pack)."""
    N, K, H, W, ldx = _g(x, 'class_scores_fwd')
    Nc = w.shape[0]
    wp, _, bpad = packed
    out = empty_nhwc(N, Nc, H, W, x.dtype, x.device, pitch=max(_pad8(Nc) + 8, wp.shape[0]))
    _lib.call('tss_pwconv_fwd', x=x, w=w, wp=wp, y=out, M=N * H * W, K=K, Nc=wp.shape[0], ldx=ldx, ldy=_g(out, 'class_scores_fwd')[4],
              scale=None, shift=bpad, res=None, ldr=0, flags=0, stats=None, impl=1, dtype=dtype_code(x.dtype))
    return out


def class_scores_dgrad(dy, w, packed):
    N, Nc, H, W, lddy = _g(dy, 'class_scores_dgrad')
    K = w.shape[1]
    wpT = packed[1]
    if lddy < wpT.shape[1]:
        raise RuntimeError('class_scores_dgrad: gradient pitch %d < packed pitch %d' % (lddy, wpT.shape[1]))
    dx = empty_nhwc(N, K, H, W, dy.dtype, dy.device)
    _lib.call('tss_pwconv_dgrad', dy=dy, w=w, wpT=wpT, dx=dx, M=N * H * W, K=K, Nc=wpT.shape[1], lddy=lddy, lddx=K, impl=1,
              dtype=dtype_code(dy.dtype))
    return dx


def dwpw_fwd(x, w_dw, scale1, shift1, relu1, wp, scale2, shift2, res=None, relu2=False):
    """Inference: act2(BN2(pw(act1(BN1(dw3x3(x))))) [+ res]) in one kernel (stride 1, bf16)."""
    N, C, H, W, ld = _g(x, 'dwpw_fwd')
    if ld != C or x.dtype != torch.bfloat16:
        raise RuntimeError('dwpw_fwd: expects a dense bf16 NHWC input')
    Nc = wp.shape[0]
    y = empty_nhwc(N, Nc, H, W, x.dtype, x.device)
    _lib.call('tss_dwpw_fwd', x=x, w_dw=w_dw, scale1=scale1, shift1=shift1, flags1=_flags(relu1), wp=wp, y=y,
              N=N, H=H, W=W, C=C, Nc=Nc, ldy=Nc, scale2=scale2, shift2=shift2, res=res,
              ldr=_g(res, 'dwpw_fwd')[4] if res is not None else 0, flags2=_flags(relu2))
    return y


def pack_weights_bf16(w, out=None):
    """-> (wp, wpT) bf16 copies of a pointwise weight; ``out``: an existing pair to refresh in place."""
    Nc, K = w.shape[0], w.shape[1]
    if out is not None and out[0].device == w.device and tuple(out[0].shape) == (Nc, K) and tuple(out[1].shape) == (K, Nc):
        wp, wpT = out
    else:
        wp = torch.empty((Nc, K), dtype=torch.bfloat16, device=w.device)
        wpT = torch.empty((K, Nc), dtype=torch.bfloat16, device=w.device)
    _lib.call('tss_pack_weights_bf16', w=w, wp=wp, wpT=wpT, Nc=Nc, K=K)
    return wp, wpT


# ------------------------------------------------------------------ stem ---------------
def stem_fwd(x, w, dtype, scale=None, shift=None, relu=False, stats=None):
    """x: NCHW fp32 contiguous (N,3,H,W) image batch, read in place."""
    N, Cin, H, W = x.shape
    if Cin != 3 or x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError('stem_fwd: expects a contiguous float32 (N,3,H,W) batch')
    Cout = w.shape[0]
    y = empty_nhwc(N, Cout, (H - 1) // 2 + 1, (W - 1) // 2 + 1, dtype, x.device)
    _lib.call('tss_stem3x3s2_fwd', x=x, w=w, y=y, N=N, H=H, W=W, Cout=Cout, scale=scale,
              shift=shift, flags=_flags(relu), stats=stats, dtype=dtype_code(dtype))
    return y


def stem_fwd_tc(x, w, scale=None, shift=None, relu=False, stats=None):
    """``stem_fwd`` on the tensor cores (bf16 output)."""
    N, Cin, H, W = x.shape
    if Cin != 3 or x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError('stem_fwd_tc: expects a contiguous float32 (N,3,H,W) batch')
    Cout = w.shape[0]
    y = empty_nhwc(N, Cout, (H - 1) // 2 + 1, (W - 1) // 2 + 1, torch.bfloat16, x.device)
    _lib.call('tss_stem3x3s2_fwd_tc', x=x, w=w, y=y, N=N, H=H, W=W, Cout=Cout, scale=scale, shift=shift,
              flags=_flags(relu), stats=stats)
    return y


def stem_wgrad_tc(x, dy, dw):
    """``stem_wgrad`` on the tensor cores (bf16 gradient, dense NHWC)."""
    N, _, H, W = x.shape
    if dy.dtype != torch.bfloat16 or _g(dy, 'stem_wgrad_tc')[4] != dy.shape[1]:
        raise RuntimeError('stem_wgrad_tc: expects a dense bfloat16 NHWC gradient')
    if gate('STEM_WGRAD_PATCHES'):
        # patch matrix + the TMA-fed pointwise weight-gradient GEMM instead of thread-built operands
        stem_wgrad_from_patches(stem_patches(x), dy, dw)
        return
    _lib.call('tss_stem3x3s2_wgrad_tc', x=x, dy=dy, dw=dw, N=N, H=H, W=W, Cout=dy.shape[1])


def stem_patches(x):
    """bf16 patch matrix (N*Ho*Wo, 32) of the stem convolution: the 27 taps of every output pixel, 5 zero columns."""
    N, _, H, W = x.shape
    M = N * ((H - 1) // 2 + 1) * ((W - 1) // 2 + 1)
    patches = torch.empty((M, 32), dtype=torch.bfloat16, device=x.device)
    _lib.call('tss_stem3x3s2_patches', x=x, patches=patches, N=N, H=H, W=W)
    return patches


def stem_wgrad_from_patches(patches, dy, dw):
    """dw (32,3,3,3) += dy^T . patches on the TMA-fed tensor-core weight-gradient GEMM of the pointwise convs."""
    if dy.dtype != torch.bfloat16 or _g(dy, 'stem_wgrad_from_patches')[4] != dy.shape[1]:
        raise RuntimeError('stem_wgrad_from_patches: expects a dense bfloat16 NHWC gradient')
    dw32 = torch.empty((32, 32), dtype=torch.float32, device=dy.device)
    _lib.call('tss_stem3x3s2_wgrad_from_patches', patches=patches, dy=dy, dw32=dw32, dw=dw, M=patches.shape[0], Cout=dy.shape[1])


def stem_wgrad_tc_bn(x, dz, y, mean, rstd, gamma, beta, sums, relu, dw, dgamma=None, dbeta=None):
    """Stem weight gradient on the tensor cores with the BatchNorm-backward apply folded in (no dy tensor)."""
    N, _, H, W = x.shape
    C = dz.shape[1]
    if dz.dtype != torch.bfloat16 or _g(dz, 'stem_wgrad_tc_bn')[4] != C or _g(y, 'stem_wgrad_tc_bn')[4] != C:
        raise RuntimeError('stem_wgrad_tc_bn: expects dense bfloat16 NHWC tensors')
    _lib.call('tss_stem3x3s2_wgrad_tc_bn', x=x, dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums,
              flags=_flags(relu), count=dz.shape[0] * dz.shape[2] * dz.shape[3], dw=dw, dgamma=dgamma, dbeta=dbeta, N=N, H=H, W=W,
              Cout=C)


def stem_wgrad(x, dy, dw):
    N, _, H, W = x.shape
    _lib.call('tss_stem3x3s2_wgrad', x=x, dy=dy, dw=dw, N=N, H=H, W=W, Cout=dy.shape[1],
              dtype=dtype_code(dy.dtype))


# ------------------------------------------------------------------ dense 3x3 ----------
def im2col3x3(x):
    """(N, C, H, W) NHWC -> tap-major patch matrix as an NHWC tensor of logical shape (N, 9C, H, W)."""
    N, C, H, W, ld = _g(x, 'im2col3x3')
    if ld != C:
        raise RuntimeError('im2col3x3: pitched input not supported')
    col = empty_nhwc(N, 9 * C, H, W, x.dtype, x.device)
    _lib.call('tss_im2col3x3', x=x, col=col, N=N, H=H, W=W, C=C, dtype=dtype_code(x.dtype))
    return col


def col2im3x3(dcol):
    N, C9, H, W, ld = _g(dcol, 'col2im3x3')
    if ld != C9 or C9 % 9:
        raise RuntimeError('col2im3x3: expects a dense (N, 9C, H, W) NHWC gradient')
    dx = empty_nhwc(N, C9 // 9, H, W, dcol.dtype, dcol.device)
    _lib.call('tss_col2im3x3', dcol=dcol, dx=dx, N=N, H=H, W=W, C=C9 // 9, dtype=dtype_code(dcol.dtype))
    return dx


def permute_weights3x3(w):
    """(Cout, Cin, 3, 3) fp32 -> tap-major (Cout, 9*Cin, 1, 1) fp32."""
    Cout, Cin = w.shape[0], w.shape[1]
    wk = torch.empty((Cout, 9 * Cin, 1, 1), dtype=torch.float32, device=w.device)
    _lib.call('tss_permute_weights3x3', src=w, dst=wk, Cout=Cout, Cin=Cin, backward=0)
    return wk


def permute_weights3x3_bwd(dwk, dw):
    """dw (Cout, Cin, 3, 3) += tap-major gradient dwk (Cout, 9*Cin, 1, 1)."""
    Cout, Cin = dw.shape[0], dw.shape[1]
    _lib.call('tss_permute_weights3x3', src=dwk, dst=dw, Cout=Cout, Cin=Cin, backward=1)
    return dw


# ------------------------------------------------------------------ batch norm ---------
def layer_scratch(bn, device):
    """Per-layer scratch of 3C doubles: [conv-epilogue statistics, 2C fp64 | BatchNorm-backward sums,
    2C fp32 in the last C doubles], allocated once and kept zero by ``tss_bn_finalize``'s
    consume-and-clear (no memset launch per layer and step).  ``bn._tss_dirty`` tracks whether the
    backward part was used since it was last cleared."""
    s = getattr(bn, '_tss_scratch', None)
    if s is None or s.device != device or s.numel() != 3 * bn.num_features:
        s = torch.zeros(3 * bn.num_features, dtype=torch.float64, device=device)
        bn._tss_scratch = s
        bn._tss_dirty = False
    return s


def bn_finalize(stats, count, bn, momentum, eps, update_running=True, clear_n=0, C=None):
    """-> scale, shift, mean, rstd (fp32 (C,)); updates bn's running stats in place."""
    C = stats.numel() // 2 if C is None else C
    out = torch.empty((4, C), dtype=torch.float32, device=stats.device)
    track = update_running and bn.running_mean is not None
    _lib.call('tss_bn_finalize', stats=stats, count=count, gamma=bn.weight, beta=bn.bias,
              running_mean=bn.running_mean if track else None,
              running_var=bn.running_var if track else None,
              num_batches_tracked=bn.num_batches_tracked if track else None,
              momentum=momentum, eps=eps, scale=out[0], shift=out[1], mean=out[2], rstd=out[3], C=C,
              clear_n=clear_n)
    if track:
        WEIGHTS_EPOCH[0] += 1
    return out[0], out[1], out[2], out[3]


def bn_finalize_apply(stats, count, bn, momentum, eps, y, res=None, relu=False, update_running=True, clear_n=0, C=None):
    """``bn_finalize`` + ``bn_apply`` as one launch -> z, mean, rstd."""
    N, Cy, H, W, ldy = _g(y, 'bn_finalize_apply')
    C = Cy if C is None else C
    out = torch.empty((2, C), dtype=torch.float32, device=stats.device)
    ticket = getattr(bn, '_tss_ticket', None)
    if ticket is None or ticket.device != stats.device:
        ticket = bn._tss_ticket = torch.zeros(1, dtype=torch.int32, device=stats.device)
    track = update_running and bn.running_mean is not None
    z = empty_nhwc(N, C, H, W, y.dtype, y.device)
    _lib.call('tss_bn_finalize_apply', stats=stats, count=count, gamma=bn.weight, beta=bn.bias,
              running_mean=bn.running_mean if track else None, running_var=bn.running_var if track else None,
              num_batches_tracked=bn.num_batches_tracked if track else None, momentum=momentum, eps=eps,
              mean=out[0], rstd=out[1], ticket=ticket, clear_n=clear_n, y=y, res=res, z=z, M=N * H * W, C=C, ldy=ldy,
              ldr=_g(res, 'bn_finalize_apply')[4] if res is not None else 0, ldz=C, flags=_flags(relu),
              dtype=dtype_code(y.dtype))
    if track:
        WEIGHTS_EPOCH[0] += 1
    return z, out[0], out[1]


def bn_fold(bn, out=None):
    """-> scale, shift of the eval-mode BatchNorm.  ``out`` ((2, C) fp32): refresh an existing pair in place, so that
    whoever holds its address (a captured CUDA graph, an address table) sees the new values."""
    C = bn.num_features
    if out is None or out.device != bn.running_mean.device or tuple(out.shape) != (2, C):
        out = torch.empty((2, C), dtype=torch.float32, device=bn.running_mean.device)
    _lib.call('tss_bn_fold', gamma=bn.weight, beta=bn.bias, running_mean=bn.running_mean,
              running_var=bn.running_var, eps=bn.eps, scale=out[0], shift=out[1], C=C)
    return out[0], out[1]


def bn_apply(y, scale, shift, y2=None, scale2=None, shift2=None, res=None, relu=False, out=None):
    N, C, H, W, ldy = _g(y, 'bn_apply')
    if out is None:
        out = empty_nhwc(N, C, H, W, y.dtype, y.device)
    _lib.call('tss_bn_apply', y=y, scale=scale, shift=shift, y2=y2, scale2=scale2, shift2=shift2,
              res=res, z=out, M=N * H * W, C=C, ldy=ldy,
              ldy2=_g(y2, 'bn_apply')[4] if y2 is not None else 0,
              ldr=_g(res, 'bn_apply')[4] if res is not None else 0,
              ldz=_g(out, 'bn_apply')[4], flags=_flags(relu), dtype=dtype_code(y.dtype))
    return out


_GRID_SYNC = {}
BN_BWD_ONEPASS = gate('BN_BWD_ONEPASS')      # module attribute: tests toggle it
_ONEPASS_MAX_BYTES = int(float(os.environ.get('TSS_BN_ONEPASS_MB', '64')) * (1 << 20))


def _grid_sync(device):
    """The barrier words of ``tss_bn_bwd_onepass`` ({arrivals, generation, sticky time-out flag, unused}): one buffer per
    device, so the BatchNorm backward passes of a device must be stream-ordered (they are: autograd runs a device's
    backward on one stream; a caller with several concurrent streams passes its own buffers through the C ABI).  It is
    allocated on first use, which must not happen inside a CUDA-graph capture (the buffer would live in that graph's
    private pool and die with it): every graphed step of this package warms up eagerly first."""
    key = device.index if device.type == 'cuda' else 'cpu'
    buf = _GRID_SYNC.get(key)
    if buf is None:
        if device.type == 'cuda' and torch.cuda.is_current_stream_capturing():
            raise RuntimeError('tss_b200: the first BatchNorm backward of a process ran inside a CUDA-graph capture; '
                               'run one eager step before capturing (or set TSS_BN_BWD_ONEPASS=0)')
        buf = _GRID_SYNC[key] = torch.zeros(4, dtype=torch.int32, device=device)
    return buf


def grid_sync_timed_out():
    """True if a grid barrier of ``tss_bn_bwd_onepass`` ever gave up waiting (its results are then wrong)."""
    return any(int(buf[2]) != 0 for buf in _GRID_SYNC.values())


def _onepass_fits(dz, y, z):
    """``tss_bn_bwd_onepass`` pays off while its operands stay in L2 between its two passes."""
    return BN_BWD_ONEPASS and sum(t.numel() * t.element_size() for t in (dz, y, z) if t is not None) <= _ONEPASS_MAX_BYTES


def bn_backward(dz, z, y, mean, rstd, gamma, relu, want_dres=False, dgamma=None, dbeta=None, beta=None,
                sums=None, sync=None, prereduced=False):
    """-> dy, dres (or None).  dgamma/dbeta (fp32 (C,)) are accumulated into.  ``z=None`` with
    ``relu``: the ReLU mask is recomputed from ``y`` (needs ``beta``; no residual before the ReLU)."""
    N, C, H, W, lddz = _g(dz, 'bn_backward')
    M = N * H * W
    if sums is None:
        sums = zeros_f32(2 * C, dz.device)
    fl = _flags(relu)
    code = dtype_code(dz.dtype)
    ldz = _g(z, 'bn_backward')[4] if z is not None else 0
    ldy = _g(y, 'bn_backward')[4]
    if prereduced:      # dz already is g = dz * mask and `sums` already holds the two sums (fused into the
        fl = 0          # dgrad epilogue of the consumer): the apply pass alone, without a mask
        relu = False
    elif (sync is None or sync[0] <= 1) and _onepass_fits(dz, y, z if relu else None):
        # both passes in one launch: dz and y (and z) stay in L2 between them
        dy = empty_nhwc(N, C, H, W, dz.dtype, dz.device)
        dres = empty_nhwc(N, C, H, W, dz.dtype, dz.device) if want_dres else None
        _lib.call('tss_bn_bwd_onepass', dz=dz, z=z if relu else None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta,
                  sums=sums, dy=dy, dres=dres, dgamma=dgamma, dbeta=dbeta, M=M, C=C, lddz=lddz, ldz=ldz, ldy=ldy, lddy=C,
                  lddres=C, flags=fl, sync=_grid_sync(dz.device), dtype=code)
        return dy, dres
    else:
        _lib.call('tss_bn_bwd_reduce', dz=dz, z=z if relu else None, y=y, mean=mean, rstd=rstd, gamma=gamma,
              beta=beta, sums=sums,
              M=M, C=C, lddz=lddz, ldz=ldz, ldy=ldy, flags=fl, dtype=code)
    world = 1
    if sync is not None and sync[0] > 1:
        # SyncBN backward: dgamma / dbeta are the LOCAL sums (the data-parallel all-reduce averages them
        # like every other gradient); the input gradient needs the sums over ALL ranks' pixels
        world = sync[0]
        if dbeta is not None:
            dbeta += sums[:C]
        if dgamma is not None:
            dgamma += sums[C:]
        dgamma = dbeta = None
        torch.distributed.all_reduce(sums, group=sync[1])
    dy = empty_nhwc(N, C, H, W, dz.dtype, dz.device)
    dres = empty_nhwc(N, C, H, W, dz.dtype, dz.device) if want_dres else None
    _lib.call('tss_bn_bwd_apply', dz=dz, z=z if relu else None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta,
              sums=sums, dy=dy, dres=dres, dgamma=dgamma, dbeta=dbeta, M=M, count=M * world, C=C, lddz=lddz, ldz=ldz,
              ldy=ldy, lddy=C, lddres=C, flags=fl, dtype=code)
    return dy, dres


def relu_bwd(dz, z):
    N, C, H, W, lddz = _g(dz, 'relu_bwd')
    g = empty_nhwc(N, C, H, W, dz.dtype, dz.device)
    _lib.call('tss_relu_bwd', dz=dz, z=z, g=g, M=N * H * W, C=C, lddz=lddz, ldz=_g(z, 'relu_bwd')[4],
              ldg=C, dtype=dtype_code(dz.dtype))
    return g


def add(a, b):
    N, C, H, W, lda = _g(a, 'add')
    out = empty_nhwc(N, C, H, W, a.dtype, a.device)
    _lib.call('tss_add', a=a, b=b, out=out, M=N * H * W, C=C, lda=lda, ldb=_g(b, 'add')[4], ldo=C,
              dtype=dtype_code(a.dtype))
    return out


def copy_rows(src, dst):
    N, C, H, W, lds = _g(src, 'copy_rows')
    _lib.call('tss_copy_rows', src=src, dst=dst, M=N * H * W, C=C, lds=lds, ldd=_g(dst, 'copy_rows')[4],
              dtype=dtype_code(src.dtype))
    return dst


def cast_from_f32(src, dtype):
    dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    _lib.call('tss_cast_from_f32', src=src, dst=dst, n=src.numel(), dtype=dtype_code(dtype))
    return dst


def scale_inplace(x, s):
    _lib.call('tss_scale_inplace', x=x, s=s, n=x.numel(), dtype=dtype_code(x.dtype))
    return x


# ------------------------------------------------------------------ pooling / resize ---
def adaptive_pool_fwd(x, bins=PYRAMID_BINS):
    """-> list of (N, C, b, b) NHWC tensors (views into one buffer), one per bin."""
    N, C, H, W, ld = _g(x, 'adaptive_pool_fwd')
    if ld != C:
        raise RuntimeError('adaptive_pool_fwd: pitched input not supported')
    cells = sum(b * b for b in bins)
    buf = torch.empty((cells * N, C), dtype=x.dtype, device=x.device)
    _lib.call('tss_adaptive_pool_fwd', x=x, out=buf, N=N, H=H, W=W, C=C, bins=_HostInts(bins),
              nbins=len(bins), dtype=dtype_code(x.dtype))
    return buf, split_pool_buffer(buf, N, C, bins)


def split_pool_buffer(buf, N, C, bins):
    outs, off = [], 0
    for b in bins:
        outs.append(buf[off:off + N * b * b].view(N, b, b, C).permute(0, 3, 1, 2))
        off += N * b * b
    return outs


def adaptive_pool_bwd(dbuf, dx, bins=PYRAMID_BINS, accumulate=False):
    N, C, H, W, ld = _g(dx, 'adaptive_pool_bwd')
    if ld != C:
        raise RuntimeError('adaptive_pool_bwd: pitched gradient not supported')
    _lib.call('tss_adaptive_pool_bwd', dout=dbuf, dx=dx, N=N, H=H, W=W, C=C, bins=_HostInts(bins),
              nbins=len(bins), accumulate=int(accumulate), dtype=dtype_code(dx.dtype))
    return dx


# ------------------------------------------------------------------ pyramid pooling, grouped ------
def ppm_branches_fwd(pool, table, N, C, Cb, bins, momentum, eps):
    """All pyramid branches in one launch: 1x1 conv + BatchNorm(train) + ReLU on the pooled rows.
    -> y (raw conv output), z (activated), mean, rstd ((nbins, Cb) fp32)."""
    rows = pool.shape[0]
    y = torch.empty((rows, Cb), dtype=pool.dtype, device=pool.device)
    z = torch.empty((rows, Cb), dtype=pool.dtype, device=pool.device)
    mean = torch.empty((len(bins), Cb), dtype=torch.float32, device=pool.device)
    rstd = torch.empty((len(bins), Cb), dtype=torch.float32, device=pool.device)
    _lib.call('tss_ppm_branches_fwd', pool=pool, table=table, y=y, z=z, mean=mean, rstd=rstd, N=N, C=C, Cb=Cb,
              bins=_HostInts(bins), nbins=len(bins), momentum=float(momentum), eps=float(eps), dtype=dtype_code(pool.dtype))
    WEIGHTS_EPOCH[0] += 1          # the kernel updated the branches' running statistics
    return y, z, mean, rstd


def ppm_branches_eval(pool, table, N, C, Cb, bins):
    """Eval mode: all pyramid branches (1x1 conv + folded BatchNorm + ReLU) in one launch -> z."""
    z = torch.empty((pool.shape[0], Cb), dtype=pool.dtype, device=pool.device)
    _lib.call('tss_ppm_branches_eval', pool=pool, table=table, z=z, N=N, C=C, Cb=Cb, bins=_HostInts(bins),
              nbins=len(bins), dtype=dtype_code(pool.dtype))
    return z


def ppm_concat_fwd(x, z, Cb, bins):
    N, C, H, W, ld = _g(x, 'ppm_concat_fwd')
    if ld != C:
        raise RuntimeError('ppm_concat_fwd: pitched input not supported')
    cat = empty_nhwc(N, C + len(bins) * Cb, H, W, x.dtype, x.device)
    _lib.call('tss_ppm_concat_fwd', x=x, z=z, cat=cat, N=N, H=H, W=W, C=C, Cb=Cb, bins=_HostInts(bins),
              nbins=len(bins), dtype=dtype_code(x.dtype))
    return cat


def ppm_concat_bwd(dcat, C, Cb, bins):
    N, Ct, H, W, ld = _g(dcat, 'ppm_concat_bwd')
    dz = torch.empty((N * sum(b * b for b in bins), Cb), dtype=dcat.dtype, device=dcat.device)
    _lib.call('tss_ppm_concat_bwd', dcat=dcat, dz=dz, N=N, H=H, W=W, C=C, Cb=Cb, lddcat=ld, bins=_HostInts(bins),
              nbins=len(bins), dtype=dtype_code(dcat.dtype))
    return dz


def ppm_branches_bwd(dz, y, pool, table, mean, rstd, N, bins):
    """-> dpool; the parameter gradients are accumulated through the addresses in ``table``."""
    C, Cb = pool.shape[1], y.shape[1]
    dy = torch.empty_like(dz)
    dpool = torch.empty_like(pool)
    _lib.call('tss_ppm_branches_bwd', dz=dz, y=y, pool=pool, table=table, mean=mean, rstd=rstd, dy=dy, dpool=dpool,
              N=N, C=C, Cb=Cb, bins=_HostInts(bins), nbins=len(bins), dtype=dtype_code(dz.dtype))
    return dpool


class _HostInts:
    """A small host int array argument (``const int*`` read by the launcher, not a kernel)."""

    def __init__(self, values):
        import ctypes
        self.values = tuple(int(v) for v in values)
        self.array = (ctypes.c_int * len(self.values))(*self.values)


def bilinear_fwd(x, Ho, Wo, out=None):
    N, C, Hi, Wi, ldx = _g(x, 'bilinear_fwd')
    if out is None:
        out = empty_nhwc(N, C, Ho, Wo, x.dtype, x.device)
    _lib.call('tss_bilinear_fwd', x=x, y=out, N=N, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, C=C, ldx=ldx,
              ldy=_g(out, 'bilinear_fwd')[4], dtype=dtype_code(x.dtype))
    return out


def bilinear_bwd(dy, Hi, Wi):
    N, C, Ho, Wo, lddy = _g(dy, 'bilinear_bwd')
    dx = empty_nhwc(N, C, Hi, Wi, dy.dtype, dy.device)
    # separable rows-then-columns transpose through an fp32 scratch (Ho/Hi + Wo/Wi taps instead
    # of their product); tiny maps take the single gather pass
    ws = None
    if N * Ho * Wo * C >= (1 << 16):
        ws = torch.empty(N * Hi * Wo * C, dtype=torch.float32, device=dy.device)
    _lib.call('tss_bilinear_bwd', dy=dy, dx=dx, workspace=ws, N=N, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, C=C,
              lddy=lddy, lddx=C, dtype=dtype_code(dy.dtype))
    return dx


def upsample_logits_fwd(x, Ho, Wo):
    """NHWC (pitched) class scores -> NCHW-contiguous (N, C, Ho, Wo)."""
    N, C, Hi, Wi, ldx = _g(x, 'upsample_logits_fwd')
    y = torch.empty((N, C, Ho, Wo), dtype=x.dtype, device=x.device)
    _lib.call('tss_upsample_logits_fwd', x=x, y=y, N=N, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, C=C, ldx=ldx,
              dtype=dtype_code(x.dtype))
    return y


def upsample_logits_bwd(dy, Hi, Wi, pitch):
    """NCHW-contiguous gradient -> NHWC gradient (N, C, Hi, Wi) with channel pitch ``pitch``."""
    N, C, Ho, Wo = dy.shape
    if not dy.is_contiguous():
        dy = dy.contiguous()
    acc = torch.zeros((N, Hi, Wi, pitch), dtype=torch.float32, device=dy.device)
    _lib.call('tss_upsample_logits_bwd', dy=dy, dx32=acc, N=N, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, C=C,
              lddx=pitch, dtype=dtype_code(dy.dtype))
    if dy.dtype != torch.float32:
        acc = cast_from_f32(acc, dy.dtype)
    return acc[..., :C].permute(0, 3, 1, 2)


def bilinear_nchw_f32(x, Ho, Wo):
    N, C, Hi, Wi = x.shape
    y = torch.empty((N, C, Ho, Wo), dtype=torch.float32, device=x.device)
    _lib.call('tss_bilinear_nchw_f32', x=x, y=y, NC=N * C, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo)
    return y


# ------------------------------------------------------------------ loss / metrics -----
_OHEM_WS = {}


def ohem_select(pixel_loss, n_keep, thresh):
    """-> loss (fp32 scalar tensor), weights (device float[4] rule for the CE kernels' ``ohem`` argument)."""
    dev = pixel_loss.device
    ws = _OHEM_WS.get(dev)
    if ws is None:          # persistent, zero-initialised; every call leaves it reusable
        ws = _OHEM_WS[dev] = torch.zeros(int(_lib.call('tss_ohem_workspace_bytes')), dtype=torch.uint8, device=dev)
    out = torch.empty(8, dtype=torch.float32, device=dev)
    flat = pixel_loss.reshape(-1)
    _lib.call('tss_ohem_select', pixel_loss=flat, n=flat.numel(), n_keep=int(n_keep), thresh=float(thresh),
              workspace=ws, loss=out[:1], weights=out[4:])
    return out[0], out[4:]


def ce_forward(logits, target, ignore_index, want_grad, want_pixel_loss=False, ohem=None):
    """-> loss (fp32 scalar tensor), dlogits (or None), pixel_loss (or None), nvalid (int64 (1,))."""
    N, C, H, W = logits.shape
    if not logits.is_contiguous():
        logits = logits.contiguous()
    if target.dtype != torch.int64 or not target.is_contiguous():
        target = target.long().contiguous()
    dev = logits.device
    nvalid = torch.empty(1, dtype=torch.int64, device=dev)
    _lib.call('tss_ce_count_valid', target=target, n=target.numel(), ignore_index=ignore_index, num_classes=C,
              nvalid=nvalid)
    loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    dlogits = torch.empty_like(logits) if want_grad else None
    pixel = torch.empty((N, H, W), dtype=torch.float32, device=dev) if want_pixel_loss else None
    _lib.call('tss_ce_fwd', logits=logits, target=target, N=N, C=C, HW=H * W, ignore_index=ignore_index,
              nvalid=nvalid, loss_sum=loss_sum, pixel_loss=pixel, dlogits=dlogits, ohem=ohem,
              dtype=dtype_code(logits.dtype))
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _lib.call('tss_ce_finalize', loss_sum=loss_sum, nvalid=nvalid, loss=loss)
    return loss, dlogits, pixel, nvalid


def upsample_ce_forward(scores, target, Ho, Wo, ignore_index, want_grad, want_pixel_loss=False, ohem=None):
    """Fused head on the 1/8-resolution NHWC class scores: -> loss (fp32 scalar tensor), gradient
    w.r.t. ``scores`` (same logical shape, channel pitch padded to 8; or None), per-pixel loss (or None)."""
    N, C, Hi, Wi, ldx = _g(scores, 'upsample_ce_forward')
    if target.dtype != torch.int64 or not target.is_contiguous():
        target = target.long().contiguous()
    if tuple(target.shape) != (N, Ho, Wo):
        raise ValueError('upsample_ce_forward: target shape %s does not match (%d, %d, %d)' % (tuple(target.shape), N, Ho, Wo))
    dev = scores.device
    pitch = _pad8(C)
    acc = torch.zeros((N, Hi, Wi, pitch), dtype=torch.float32, device=dev) if want_grad else None
    # [loss_sum (fp64) | nvalid (int64 bits)] zeroed by one fill
    red = torch.zeros(2, dtype=torch.float64, device=dev)
    nvalid = red[1:].view(torch.int64)
    pixel = torch.empty((N, Ho, Wo), dtype=torch.float32, device=dev) if want_pixel_loss else None
    code = dtype_code(scores.dtype)
    _lib.call('tss_upsample_ce_fwd', x=scores, target=target, N=N, C=C, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, ldx=ldx,
              ignore_index=ignore_index, loss_sum=red[:1], nvalid=nvalid, pixel_loss=pixel, dx32=acc, lddx=pitch,
              ohem=ohem, dtype=code)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dx = torch.empty((N, Hi, Wi, pitch), dtype=scores.dtype, device=dev) if want_grad else None
    if ohem is not None:       # OHEM-weighted gradient: already carries its per-pixel weights
        nvalid = torch.ones(1, dtype=torch.int64, device=dev)
    _lib.call('tss_upsample_ce_finalize', loss_sum=red[:1], nvalid=nvalid, loss=loss, dx32=acc, dx=dx,
              n=N * Hi * Wi * pitch if want_grad else 0, dtype=code)
    if dx is not None:
        dx = dx[..., :C].permute(0, 3, 1, 2)
    return loss, dx, pixel


def confusion_from_labels(pred, target, num_classes, cm):
    """cm (int64 (C,C), device) += histogram of (target, pred)."""
    pred = pred.long().contiguous().view(-1)
    target = target.long().contiguous().view(-1)
    _lib.call('tss_confusion_from_labels', pred=pred, target=target, n=pred.numel(), C=num_classes, cm=cm)
    return cm


def confusion_from_logits(logits, target, cm, want_pred=False):
    N, C, H, W = logits.shape
    if not logits.is_contiguous():
        logits = logits.contiguous()
    target = target.long().contiguous()
    pred = torch.empty((N, H, W), dtype=torch.int64, device=logits.device) if want_pred else None
    _lib.call('tss_confusion_from_logits', logits=logits, target=target, N=N, C=C, HW=H * W, cm=cm,
              pred_out=pred, dtype=dtype_code(logits.dtype))
    return pred


def adamw_step(p, g, m, v, hyper, grad_scale=1.0):
    _lib.call('tss_adamw_step', p=p, g=g, m=m, v=v, n=p.numel(), hyper=hyper, grad_scale=float(grad_scale))
    WEIGHTS_EPOCH[0] += 1


# ------------------------------------------------------------------ input pipeline ------
class _HostFloats:
    """A small host float array argument (``const float*`` read by the launcher, not a kernel)."""

    def __init__(self, values):
        import ctypes
        self.values = tuple(float(v) for v in values)
        self.array = (ctypes.c_float * len(self.values))(*self.values)


def augment_batch(images, labels, geom, lut, norm, crop):
    """uint8 (N,H,W,3) frames + uint8 (N,H,W) label ids + per-sample draws ``geom`` (int32 (N,5) on the device)
    -> fp32 (N,3,ch,cw) normalised crops, int64 (N,ch,cw) train ids.  ``norm``: 6 host floats."""
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3 or not images.is_contiguous():
        raise RuntimeError('augment_batch: images must be a contiguous uint8 (N,H,W,3) tensor')
    N, H, W, _ = images.shape
    if labels is not None and (labels.dtype != torch.uint8 or tuple(labels.shape) != (N, H, W) or not labels.is_contiguous()):
        raise RuntimeError('augment_batch: labels must be a contiguous uint8 (N,H,W) tensor')
    if geom.dtype != torch.int32 or tuple(geom.shape) != (N, 5) or not geom.is_contiguous():
        raise RuntimeError('augment_batch: geom must be a contiguous int32 (N,5) tensor')
    ch, cw = crop
    x = torch.empty((N, 3, ch, cw), dtype=torch.float32, device=images.device)
    y = torch.empty((N, ch, cw), dtype=torch.int64, device=images.device) if labels is not None else None
    _lib.call('tss_augment_batch', images=images, labels=labels, geom=geom, lut=lut, norm=_HostFloats(norm),
              out_image=x, out_label=y, N=N, H=H, W=W, ch=ch, cw=cw)
    return x, y


# ------------------------------------------------------------------ dropout -------------
_RNG = {}


def rng_state(device, seed=None):
    """Per-device {seed, offset, ticket} of the library's counter-based generator (int64[3]); the seed comes from
    torch's generator on first use (``torch.manual_seed`` before the first training forward keeps runs repeatable)."""
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:       # 'cuda' and 'cuda:<current>' are one generator
        device = torch.device('cuda', torch.cuda.current_device())
    st = _RNG.get(device)
    if st is None or seed is not None:
        s = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        st = _RNG[device] = torch.tensor([s, 0, 0], dtype=torch.int64).to(device)
    return st


def dropout_fwd(x, p):
    """-> y, used (the offset the mask was drawn with, int64[1] on the device; keep it for the backward pass)."""
    N, C, H, W, ld = _g(x, 'dropout_fwd')
    if ld != C:
        raise RuntimeError('dropout_fwd: pitched tensors not supported')
    y = empty_nhwc(N, C, H, W, x.dtype, x.device)
    used = torch.empty(1, dtype=torch.int64, device=x.device)
    _lib.call('tss_dropout_fwd', x=x, y=y, n=x.numel(), p=float(p), rng=rng_state(x.device), used=used,
              dtype=dtype_code(x.dtype))
    return y, used


def dropout_bwd(dy, p, used):
    N, C, H, W, ld = _g(dy, 'dropout_bwd')
    if ld != C:
        raise RuntimeError('dropout_bwd: pitched tensors not supported')
    dx = empty_nhwc(N, C, H, W, dy.dtype, dy.device)
    _lib.call('tss_dropout_bwd', dy=dy, dx=dx, n=dy.numel(), p=float(p), rng=rng_state(dy.device), used=used,
              dtype=dtype_code(dy.dtype))
    return dx
