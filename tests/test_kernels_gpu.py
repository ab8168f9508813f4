"""Kernel parity on the GPU: every C-ABI entry point is called with CUDA tensors and compared
with stock torch fp32 ops on the CPU (tests/fake_backend.py doubles as the per-kernel
reference), on identical seeded inputs, at the Fast-SCNN / ContextNet shapes (SURVEY.md
Appendix E, spatially reduced) plus ragged / odd shapes.

Tolerances: fp32 activations 1e-4 relative (L2) -- observed ~1e-6; bf16 2e-2 relative with the
inputs rounded to bf16 on both sides (so only accumulation order and the final rounding differ);
integer results (confusion matrix, argmax, counts) bit-exact.
"""
import math

import numpy as np
import pytest
import torch

from oracle import confusion as o_cm
from tests.fake_backend import FakeBackend
from torch_semantic_segmentation_b200 import _lib, ops

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def pair(N, C, H, W, dtype, g, pitch=None, scale=1.0):
    """(cpu, cuda) logical-NCHW / physical-NHWC tensors with identical contents and layout."""
    ld = C if pitch is None else pitch
    base = (torch.randn(N, H, W, ld, generator=g) * scale).to(dtype)
    gb = base.cuda()
    return base[..., :C].permute(0, 3, 1, 2), gb[..., :C].permute(0, 3, 1, 2)


def both(name, kc, kg):
    """Run entry point `name` on the CPU emulation (kc) and on the GPU through the C ABI (kg)."""
    FakeBackend().call(name, kc)
    _lib.backend().call(name, kg)
    torch.cuda.synchronize()


def dev(t):
    return None if t is None else t.cuda()


# ------------------------------------------------------------------ depthwise 3x3 ------
DW_CASES = [  # C, stride, dilation, N, H, W
    (32, 2, 1, 2, 40, 56), (48, 2, 1, 2, 20, 28), (384, 2, 1, 2, 16, 24), (384, 1, 1, 2, 12, 16),
    (576, 1, 1, 2, 6, 8), (768, 1, 1, 2, 6, 8), (128, 1, 4, 2, 24, 40), (128, 1, 1, 2, 24, 40),
    (192, 1, 1, 1, 9, 13), (288, 2, 1, 1, 9, 13), (64, 2, 1, 3, 7, 5), (8, 1, 1, 1, 1, 1), (40, 1, 1, 1, 5, 3),
]


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,s,d,N,H,W', DW_CASES)
def test_dwconv(C, s, d, N, H, W, dtype):
    g = gen(C + H)
    xc, xg = pair(N, C, H, W, dtype, g)
    w = torch.randn(C, 1, 3, 3, generator=g) * 0.3
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    code = _lib.dtype_code(dtype)
    # training form: raw output + BatchNorm statistics
    yc, yg = pair(N, C, Ho, Wo, dtype, g)
    sc, sg = torch.zeros(2 * C, dtype=torch.float64), torch.zeros(2 * C, dtype=torch.float64).cuda()
    common = dict(N=N, Hi=H, Wi=W, C=C, stride=s, dilation=d, dtype=code)
    both('tss_dwconv3x3_fwd', dict(x=xc, w=w, y=yc, scale=None, shift=None, flags=0, stats=sc, **common),
         dict(x=xg, w=w.cuda(), y=yg, scale=None, shift=None, flags=0, stats=sg, **common))
    assert rel(yg, yc) < TOL[dtype]
    assert rel(sg, sc) < 1e-4
    # inference form: folded BatchNorm + ReLU in the epilogue
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    both('tss_dwconv3x3_fwd', dict(x=xc, w=w, y=yc, scale=scale, shift=shift, flags=1, stats=None, **common),
         dict(x=xg, w=w.cuda(), y=yg, scale=scale.cuda(), shift=shift.cuda(), flags=1, stats=None, **common))
    assert rel(yg, yc) < TOL[dtype]
    # dgrad / wgrad
    dyc, dyg = pair(N, C, Ho, Wo, dtype, g)
    dxc, dxg = pair(N, C, H, W, dtype, g)
    both('tss_dwconv3x3_dgrad', dict(dy=dyc, w=w, dx=dxc, **common), dict(dy=dyg, w=w.cuda(), dx=dxg, **common))
    assert rel(dxg, dxc) < TOL[dtype]
    dwc = torch.randn(C, 1, 3, 3, generator=g)         # accumulate semantics
    dwg = dwc.clone().cuda()
    both('tss_dwconv3x3_wgrad', dict(x=xc, dy=dyc, dw=dwc, **common), dict(x=xg, dy=dyg, dw=dwg, **common))
    assert rel(dwg, dwc) < 1e-4


# ------------------------------------------------------------------ pointwise 1x1 ------
PW_CASES = [  # K, Nc, N, H, W
    (32, 48, 2, 20, 28), (48, 64, 2, 10, 14), (64, 384, 2, 10, 14), (384, 64, 2, 5, 7), (384, 96, 2, 3, 4),
    (96, 576, 2, 3, 4), (576, 128, 2, 3, 4), (128, 768, 2, 3, 4), (768, 128, 2, 3, 4), (128, 32, 2, 1, 1),
    (256, 128, 2, 3, 4), (128, 128, 2, 12, 20), (64, 128, 1, 12, 20), (128, 19, 2, 12, 20), (192, 32, 1, 9, 7),
    (288, 48, 1, 9, 7), (32, 32, 1, 17, 3),
]


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('K,Nc,N,H,W', PW_CASES)
def test_pwconv_simt(K, Nc, N, H, W, dtype):
    g = gen(K * 7 + Nc)
    code = _lib.dtype_code(dtype)
    M = N * H * W
    xc, xg = pair(N, K, H, W, dtype, g)
    w = torch.randn(Nc, K, 1, 1, generator=g) / math.sqrt(K)
    ldy = Nc if Nc % 8 == 0 else (Nc + 7) // 8 * 8 + 8
    yc, yg = pair(N, Nc, H, W, dtype, g, pitch=ldy)
    sc, sg = torch.zeros(2 * Nc, dtype=torch.float64), torch.zeros(2 * Nc, dtype=torch.float64).cuda()
    base = dict(M=M, K=K, Nc=Nc, ldx=K, ldy=ldy, ldr=0, impl=0, dtype=code, wp=None)
    both('tss_pwconv_fwd', dict(x=xc, w=w, y=yc, scale=None, shift=None, res=None, flags=0, stats=sc, **base),
         dict(x=xg, w=w.cuda(), y=yg, scale=None, shift=None, res=None, flags=0, stats=sg, **base))
    assert rel(yg, yc) < TOL[dtype]
    assert rel(sg, sc) < 1e-4
    if Nc % 8 == 0:
        rc, rg = pair(N, Nc, H, W, dtype, g)
        scale, shift = torch.rand(Nc, generator=g) + 0.5, torch.randn(Nc, generator=g)
        base['ldr'] = Nc
        both('tss_pwconv_fwd', dict(x=xc, w=w, y=yc, scale=scale, shift=shift, res=rc, flags=1, stats=None, **base),
             dict(x=xg, w=w.cuda(), y=yg, scale=scale.cuda(), shift=shift.cuda(), res=rg, flags=1, stats=None, **base))
        assert rel(yg, yc) < TOL[dtype]
    else:
        bias = torch.randn(Nc, generator=g)
        both('tss_pwconv_fwd', dict(x=xc, w=w, y=yc, scale=None, shift=bias, res=None, flags=0, stats=None, **base),
             dict(x=xg, w=w.cuda(), y=yg, scale=None, shift=bias.cuda(), res=None, flags=0, stats=None, **base))
        assert rel(yg, yc) < TOL[dtype]
        pad = yg.permute(0, 2, 3, 1)      # pad columns written by the kernel must be zero
        full = torch.as_strided(pad, (N, H, W, ldy), (H * W * ldy, W * ldy, ldy, 1))
        assert float(full[..., Nc:(Nc + 7) // 8 * 8].abs().max()) == 0.0
    # backward: gradient buffers with zeroed pad columns
    dbase_c = torch.zeros(N, H, W, ldy, dtype=dtype)
    dbase_c[..., :Nc] = torch.randn(N, H, W, Nc, generator=g).to(dtype)
    dyc, dyg = dbase_c[..., :Nc].permute(0, 3, 1, 2), dbase_c.cuda()[..., :Nc].permute(0, 3, 1, 2)
    dxc, dxg = pair(N, K, H, W, dtype, g)
    both('tss_pwconv_dgrad', dict(dy=dyc, w=w, wpT=None, dx=dxc, M=M, K=K, Nc=Nc, lddy=ldy, lddx=K, impl=0, dtype=code),
         dict(dy=dyg, w=w.cuda(), wpT=None, dx=dxg, M=M, K=K, Nc=Nc, lddy=ldy, lddx=K, impl=0, dtype=code))
    assert rel(dxg, dxc) < TOL[dtype]
    dwc = torch.randn(Nc, K, 1, 1, generator=g)
    dbc = torch.randn(Nc, generator=g)
    dwg, dbg = dwc.clone().cuda(), dbc.clone().cuda()
    both('tss_pwconv_wgrad', dict(x=xc, dy=dyc, dw=dwc, db=dbc, M=M, K=K, Nc=Nc, ldx=K, lddy=ldy, impl=0, dtype=code),
         dict(x=xg, dy=dyg, dw=dwg, db=dbg, M=M, K=K, Nc=Nc, ldx=K, lddy=ldy, impl=0, dtype=code))
    assert rel(dwg, dwc) < 1e-4 and rel(dbg, dbc) < 1e-4


def test_pwconv_pitched_input_from_concat_buffer():
    g = gen(5)
    catc, catg = pair(2, 256, 3, 5, torch.float32, g)
    w = torch.randn(32, 128, 1, 1, generator=g) * 0.1
    yc, yg = pair(2, 32, 3, 5, torch.float32, g)
    kw = dict(M=30, K=128, Nc=32, ldx=256, ldy=32, ldr=0, impl=0, dtype=0, wp=None, scale=None, shift=None, res=None, flags=0, stats=None)
    both('tss_pwconv_fwd', dict(x=catc[:, 128:], w=w, y=yc, **kw), dict(x=catg[:, 128:], w=w.cuda(), y=yg, **kw))
    assert rel(yg, yc) < 1e-5


# ------------------------------------------------------------------ stem ---------------
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,H,W', [(2, 64, 96), (1, 32, 32), (3, 34, 70)])
def test_stem(N, H, W, dtype):
    g = gen(H)
    code = _lib.dtype_code(dtype)
    x = torch.randn(N, 3, H, W, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    yc, yg = pair(N, 32, Ho, Wo, dtype, g)
    sc, sg = torch.zeros(64, dtype=torch.float64), torch.zeros(64, dtype=torch.float64).cuda()
    kw = dict(N=N, H=H, W=W, Cout=32, dtype=code)
    both('tss_stem3x3s2_fwd', dict(x=x, w=w, y=yc, scale=None, shift=None, flags=0, stats=sc, **kw),
         dict(x=x.cuda(), w=w.cuda(), y=yg, scale=None, shift=None, flags=0, stats=sg, **kw))
    assert rel(yg, yc) < TOL[dtype] and rel(sg, sc) < 1e-4
    scale, shift = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g)
    both('tss_stem3x3s2_fwd', dict(x=x, w=w, y=yc, scale=scale, shift=shift, flags=1, stats=None, **kw),
         dict(x=x.cuda(), w=w.cuda(), y=yg, scale=scale.cuda(), shift=shift.cuda(), flags=1, stats=None, **kw))
    assert rel(yg, yc) < TOL[dtype]
    dyc, dyg = pair(N, 32, Ho, Wo, dtype, g)
    dwc = torch.randn(32, 3, 3, 3, generator=g)
    dwg = dwc.clone().cuda()
    both('tss_stem3x3s2_wgrad', dict(x=x, dy=dyc, dw=dwc, **kw), dict(x=x.cuda(), dy=dyg, dw=dwg, **kw))
    assert rel(dwg, dwc) < 1e-4


# ------------------------------------------------------------------ batch norm ---------
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,N,H,W', [(32, 2, 20, 28), (384, 2, 5, 7), (48, 3, 9, 5), (768, 2, 3, 4), (128, 4, 1, 1)])
def test_batchnorm_family(C, N, H, W, dtype):
    g = gen(C + N)
    code = _lib.dtype_code(dtype)
    M = N * H * W
    yc, yg = pair(N, C, H, W, dtype, g)
    # finalize
    stats = torch.cat([torch.stack([yc.double().sum((0, 2, 3)), (yc.double() ** 2).sum((0, 2, 3))]).reshape(-1),
                       torch.ones(C, dtype=torch.float64)])   # [statistics | the backward-sums part that must get cleared]
    stats_g = stats.cuda()
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    outs_c = [torch.empty(C) for _ in range(4)]
    outs_g = [torch.empty(C).cuda() for _ in range(4)]
    rm, rv, nbt = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5, torch.tensor(3)
    rmg, rvg, nbtg = rm.clone().cuda(), rv.clone().cuda(), nbt.clone().cuda()
    names = ('scale', 'shift', 'mean', 'rstd')
    both('tss_bn_finalize', dict(stats=stats, count=M, gamma=gamma, beta=beta, running_mean=rm, running_var=rv,
                                 num_batches_tracked=nbt, momentum=0.1, eps=1e-5, C=C, clear_n=3 * C, **dict(zip(names, outs_c))),
         dict(stats=stats_g, clear_n=3 * C, count=M, gamma=gamma.cuda(), beta=beta.cuda(), running_mean=rmg, running_var=rvg,
              num_batches_tracked=nbtg, momentum=0.1, eps=1e-5, C=C, **dict(zip(names, outs_g))))
    for a, b in zip(outs_g, outs_c):
        assert rel(a, b) < 1e-5
    assert not stats_g.any() and not stats.any()        # consume-and-clear
    assert rel(rmg, rm) < 1e-6 and rel(rvg, rv) < 1e-6 and int(nbtg) == 4
    # against nn.functional.batch_norm's own statistics
    want_var = yc.float().var((0, 2, 3), unbiased=False)
    assert rel(outs_g[3], 1 / torch.sqrt(want_var + 1e-5)) < 1e-4
    fc, fg = [torch.empty(C), torch.empty(C)], [torch.empty(C).cuda(), torch.empty(C).cuda()]
    both('tss_bn_fold', dict(gamma=gamma, beta=beta, running_mean=rm, running_var=rv, eps=1e-5, scale=fc[0], shift=fc[1], C=C),
         dict(gamma=gamma.cuda(), beta=beta.cuda(), running_mean=rmg, running_var=rvg, eps=1e-5, scale=fg[0], shift=fg[1], C=C))
    assert rel(fg[0], fc[0]) < 1e-6 and rel(fg[1], fc[1]) < 1e-6
    # apply (+ second branch, + residual, + relu)
    scale, shift, mean, rstd = outs_c
    y2c, y2g = pair(N, C, H, W, dtype, g)
    rc, rg = pair(N, C, H, W, dtype, g)
    zc, zg = pair(N, C, H, W, dtype, g)
    kw = dict(M=M, C=C, ldy=C, ldy2=C, ldr=C, ldz=C, dtype=code)
    for use2, user, relu in [(False, False, 0), (False, True, 1), (True, False, 1), (False, False, 1)]:
        both('tss_bn_apply', dict(y=yc, scale=scale, shift=shift, y2=y2c if use2 else None, scale2=gamma if use2 else None,
                                  shift2=beta if use2 else None, res=rc if user else None, z=zc, flags=relu, **kw),
             dict(y=yg, scale=scale.cuda(), shift=shift.cuda(), y2=y2g if use2 else None, scale2=gamma.cuda() if use2 else None,
                  shift2=beta.cuda() if use2 else None, res=rg if user else None, z=zg, flags=relu, **kw))
        assert rel(zg, zc) < TOL[dtype]
    # backward (z from the last apply: relu, no residual)
    dzc, dzg = pair(N, C, H, W, dtype, g)
    ref_sums = None
    for relu, use_z in ((1, True), (1, False), (0, False)):      # use_z False + relu: mask recomputed from y
        sums_c, sums_g = torch.zeros(2 * C), torch.zeros(2 * C).cuda()
        kb = dict(M=M, C=C, lddz=C, ldz=C, ldy=C, flags=relu, dtype=code)
        both('tss_bn_bwd_reduce', dict(dz=dzc, z=zc if use_z else None, y=yc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums_c, **kb),
             dict(dz=dzg, z=zg if use_z else None, y=yg, mean=mean.cuda(), rstd=rstd.cuda(), gamma=gamma.cuda(), beta=beta.cuda(), sums=sums_g, **kb))
        assert rel(sums_g, sums_c) < 1e-4
        if relu and use_z:
            ref_sums = sums_g.clone()
        elif relu:
            assert torch.equal(sums_g.cpu() != 0, ref_sums.cpu() != 0) and rel(sums_g, ref_sums) < 1e-5   # same mask
        dyc, dyg = pair(N, C, H, W, dtype, g)
        drc, drg = pair(N, C, H, W, dtype, g)
        dgc, dbc = torch.randn(C, generator=g), torch.randn(C, generator=g)
        dgg, dbg = dgc.clone().cuda(), dbc.clone().cuda()
        both('tss_bn_bwd_apply', dict(dz=dzc, z=zc if use_z else None, y=yc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums_c,
                                      dy=dyc, dres=drc, dgamma=dgc, dbeta=dbc, lddy=C, lddres=C, count=0, **kb),
             dict(dz=dzg, z=zg if use_z else None, y=yg, mean=mean.cuda(), rstd=rstd.cuda(), gamma=gamma.cuda(), beta=beta.cuda(), sums=sums_c.cuda(),
                  dy=dyg, dres=drg, dgamma=dgg, dbeta=dbg, lddy=C, lddres=C, count=0, **kb))
        assert rel(dyg, dyc) < TOL[dtype] and rel(drg, drc) < TOL[dtype]
        assert rel(dgg, dgc) < 1e-5 and rel(dbg, dbc) < 1e-5


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,N,H,W,mask,pitch', [(128, 12, 96, 96, 'y', None), (64, 12, 48, 48, 'z', None), (96, 12, 24, 24, 'z', 104),
                                                (128, 12, 24, 24, 'y', None), (48, 3, 9, 5, 'y', 56), (768, 2, 3, 4, 'none', None),
                                                (8, 1, 1, 1, 'z', None), (384, 2, 40, 40, 'y', None), (64, 12, 96, 96, 'none', None)])
def test_batchnorm_backward_in_one_launch(C, N, H, W, mask, pitch, dtype):
    """tss_bn_bwd_onepass (pass 1 -> grid barrier -> pass 2, up to 296 resident CTAs) against the reduce + apply pair of
    the torch emulation AND against this library's own two launches; launched several times back to back (the barrier
    re-arms itself) and replayed from a CUDA graph with a dependent launch behind it; the time-out flag must stay clear."""
    g = gen(C + H)
    code = _lib.dtype_code(dtype)
    M, ld = N * H * W, C if pitch is None else pitch
    relu, use_z = int(mask != 'none'), mask == 'z'
    dzc, dzg = pair(N, C, H, W, dtype, g, pitch)
    yc, yg = pair(N, C, H, W, dtype, g, pitch)
    zc, zg = pair(N, C, H, W, dtype, g, pitch)
    if not use_z:
        zc = zg = None
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    par_g = dict(mean=mean.cuda(), rstd=rstd.cuda(), gamma=gamma.cuda(), beta=beta.cuda())
    new = lambda: torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
    sums_c, dgc, dbc, dyc, drc = torch.zeros(2 * C), torch.ones(C), torch.ones(C), new(), new() if use_z else None
    geo = dict(M=M, C=C, lddz=ld, ldz=ld if use_z else 0, ldy=ld, flags=relu, dtype=code)
    FakeBackend().call('tss_bn_bwd_onepass', dict(dz=dzc, z=zc, y=yc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums_c,
                                                  dy=dyc, dres=drc, dgamma=dgc, dbeta=dbc, lddy=C, lddres=C,
                                                  sync=torch.zeros(4, dtype=torch.int32), **geo))
    # this library's two launches
    sums_2, dy2 = torch.zeros(2 * C).cuda(), new().cuda()
    _lib.backend().call('tss_bn_bwd_reduce', dict(dz=dzg, z=zg, y=yg, sums=sums_2, **par_g, **geo))
    _lib.backend().call('tss_bn_bwd_apply', dict(dz=dzg, z=zg, y=yg, sums=sums_2, dy=dy2, dres=None, dgamma=None, dbeta=None,
                                                 lddy=C, lddres=C, count=0, **par_g, **geo))
    sync = torch.zeros(4, dtype=torch.int32).cuda()
    sums_g, dgg, dbg, dyg, drg = torch.zeros(2 * C).cuda(), torch.ones(C).cuda(), torch.ones(C).cuda(), new().cuda(), new().cuda() if use_z else None
    args = dict(dz=dzg, z=zg, y=yg, sums=sums_g, dy=dyg, dres=drg, dgamma=dgg, dbeta=dbg, lddy=C, lddres=C, sync=sync, **geo, **par_g)

    def check(tag):
        torch.cuda.synchronize()
        state = sync.cpu().tolist()
        assert state[2] == 0 and state[0] == 0, (tag, state)
        assert rel(sums_g, sums_c) < 1e-4, (tag, rel(sums_g, sums_c))
        assert rel(dyg, dyc) < TOL[dtype], (tag, rel(dyg, dyc))
        assert rel(dyg, dy2) < (1e-5 if dtype == torch.float32 else 1e-2), (tag, rel(dyg, dy2))
        assert rel(dgg, dgc) < 1e-4 and rel(dbg, dbc) < 1e-4, tag
        if use_z:
            assert torch.equal(drg.cpu(), drc), tag

    def reset():
        sums_g.zero_(); dgg.fill_(1.0); dbg.fill_(1.0); dyg.zero_()

    for i in range(3):
        reset()
        _lib.backend().call('tss_bn_bwd_onepass', args)
        check('launch %d' % i)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            reset()
            _lib.backend().call('tss_bn_bwd_onepass', args)
            _lib.backend().call('tss_add', dict(a=dyg, b=None, out=dy2, M=M, C=C, lda=C, ldb=0, ldo=C, dtype=code))   # a dependent launch
    torch.cuda.current_stream().wait_stream(st)
    for i in range(3):
        graph.replay()
        check('replay %d' % i)
    assert torch.equal(dy2, dyg)


@pytest.mark.parametrize('dtype', DTYPES)
def test_elementwise_helpers(dtype):
    g = gen(2)
    code = _lib.dtype_code(dtype)
    ac, ag = pair(2, 64, 5, 7, dtype, g, pitch=96)
    bc, bg = pair(2, 64, 5, 7, dtype, g)
    oc, og = pair(2, 64, 5, 7, dtype, g)
    kw = dict(M=70, C=64, dtype=code)
    both('tss_add', dict(a=ac, b=bc, out=oc, lda=96, ldb=64, ldo=64, **kw), dict(a=ag, b=bg, out=og, lda=96, ldb=64, ldo=64, **kw))
    assert rel(og, oc) < TOL[dtype]
    both('tss_relu_bwd', dict(dz=ac, z=bc, g=oc, lddz=96, ldz=64, ldg=64, **kw), dict(dz=ag, z=bg, g=og, lddz=96, ldz=64, ldg=64, **kw))
    assert rel(og, oc) == 0.0
    both('tss_copy_rows', dict(src=bc, dst=ac, lds=64, ldd=96, **kw), dict(src=bg, dst=ag, lds=64, ldd=96, **kw))
    assert rel(ag, ac) == 0.0
    src = torch.randn(4096, generator=g)
    dc, dg = torch.empty(4096, dtype=dtype), torch.empty(4096, dtype=dtype).cuda()
    both('tss_cast_from_f32', dict(src=src, dst=dc, n=4096, dtype=code), dict(src=src.cuda(), dst=dg, n=4096, dtype=code))
    assert rel(dg, dc) == 0.0
    s = torch.tensor([0.4])
    both('tss_scale_inplace', dict(x=dc, s=s, n=4096, dtype=code), dict(x=dg, s=s.cuda(), n=4096, dtype=code))
    assert rel(dg, dc) < TOL[dtype]


# ------------------------------------------------------------------ pooling / resize ---
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,H,W', [(2, 24, 24), (1, 32, 64), (2, 5, 7), (3, 1, 2)])
def test_adaptive_pool(N, H, W, dtype):
    g = gen(H * W)
    C, bins = 128, (1, 2, 3, 6)
    code = _lib.dtype_code(dtype)
    xc, xg = pair(N, C, H, W, dtype, g)
    cells = sum(b * b for b in bins)
    oc, og = torch.empty(cells * N, C, dtype=dtype), torch.empty(cells * N, C, dtype=dtype).cuda()
    hb = ops._HostInts(bins)
    kw = dict(N=N, H=H, W=W, C=C, bins=hb, nbins=4, dtype=code)
    both('tss_adaptive_pool_fwd', dict(x=xc, out=oc, **kw), dict(x=xg, out=og, **kw))
    assert rel(og, oc) < TOL[dtype]
    dc = torch.randn(cells * N, C, generator=g).to(dtype)
    dxc, dxg = pair(N, C, H, W, dtype, g)
    for acc in (0, 1):
        both('tss_adaptive_pool_bwd', dict(dout=dc, dx=dxc, accumulate=acc, **kw), dict(dout=dc.cuda(), dx=dxg, accumulate=acc, **kw))
        assert rel(dxg, dxc) < TOL[dtype]


RESIZE = [(2, 32, 6, 6, 24, 24), (2, 32, 1, 1, 24, 24), (1, 32, 3, 3, 32, 64), (2, 128, 6, 8, 24, 32), (1, 128, 5, 7, 20, 28),
          (1, 16, 2, 2, 5, 7), (1, 8, 7, 9, 7, 9), (1, 8, 12, 16, 5, 6)]


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,C,Hi,Wi,Ho,Wo', RESIZE)
def test_bilinear_nhwc(N, C, Hi, Wi, Ho, Wo, dtype):
    g = gen(Hi * Wo)
    code = _lib.dtype_code(dtype)
    xc, xg = pair(N, C, Hi, Wi, dtype, g)
    yc, yg = pair(N, C, Ho, Wo, dtype, g, pitch=C + 32)       # e.g. a slice of the concat buffer
    kw = dict(N=N, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, C=C, dtype=code)
    both('tss_bilinear_fwd', dict(x=xc, y=yc, ldx=C, ldy=C + 32, **kw), dict(x=xg, y=yg, ldx=C, ldy=C + 32, **kw))
    assert rel(yg, yc) < TOL[dtype]
    dxc, dxg = pair(N, C, Hi, Wi, dtype, g)
    for ws in (None, torch.empty(N * Hi * Wo * C).cuda()):       # single gather pass / separable two-pass
        both('tss_bilinear_bwd', dict(dy=yc, dx=dxc, workspace=None, lddy=C + 32, lddx=C, **kw),
             dict(dy=yg, dx=dxg, workspace=ws, lddy=C + 32, lddx=C, **kw))
        assert rel(dxg, dxc) < TOL[dtype]


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,C,Hi,Wi,s', [(2, 19, 12, 20, 8), (1, 19, 4, 4, 8), (2, 19, 3, 5, 8), (1, 11, 6, 7, 8), (1, 19, 1, 1, 8)])
def test_upsample_logits(N, C, Hi, Wi, s, dtype):
    g = gen(Hi + Wi)
    code = _lib.dtype_code(dtype)
    Ho, Wo = Hi * s, Wi * s
    xc, xg = pair(N, C, Hi, Wi, dtype, g, pitch=32)
    yc = torch.empty(N, C, Ho, Wo, dtype=dtype)
    yg = torch.empty(N, C, Ho, Wo, dtype=dtype).cuda()
    kw = dict(N=N, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, C=C, dtype=code)
    both('tss_upsample_logits_fwd', dict(x=xc, y=yc, ldx=32, **kw), dict(x=xg, y=yg, ldx=32, **kw))
    assert rel(yg, yc) < TOL[dtype]
    dy = torch.randn(N, C, Ho, Wo, generator=g).to(dtype)
    ac, ag = torch.zeros(N, Hi, Wi, 32), torch.zeros(N, Hi, Wi, 32).cuda()
    both('tss_upsample_logits_bwd', dict(dy=dy, dx32=ac, lddx=32, **kw), dict(dy=dy.cuda(), dx32=ag, lddx=32, **kw))
    assert rel(ag, ac) < 1e-4
    assert float(ag[..., C:].abs().max()) == 0.0


def test_bilinear_nchw_shrink():
    g = gen(9)
    x = torch.randn(2, 3, 64, 96, generator=g)
    for Ho, Wo in [(16, 24), (32, 48), (8, 12)]:
        yc, yg = torch.empty(2, 3, Ho, Wo), torch.empty(2, 3, Ho, Wo).cuda()
        kw = dict(NC=6, Hi=64, Wi=96, Ho=Ho, Wo=Wo)
        both('tss_bilinear_nchw_f32', dict(x=x, y=yc, **kw), dict(x=x.cuda(), y=yg, **kw))
        assert rel(yg, yc) < 1e-6


# ------------------------------------------------------------------ loss ---------------
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,H,W,frac', [(2, 24, 32, 0.1), (1, 8, 4, 0.0), (3, 5, 12, 0.5), (1, 16, 16, 1.0)])
def test_cross_entropy(N, H, W, frac, dtype):
    g = gen(H + W)
    logits = (torch.randn(N, 19, H, W, generator=g) * 3).to(dtype)
    target = torch.randint(0, 19, (N, H, W), generator=g)
    target[torch.rand(N, H, W, generator=g) < frac] = 255
    lg = logits.cuda().requires_grad_(True)
    loss, dlogits, pixel, nvalid = ops.ce_forward(lg, target.cuda(), 255, want_grad=True, want_pixel_loss=True)
    torch.cuda.synchronize()
    ref_in = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, target, ignore_index=255)
    assert int(nvalid) == int((target != 255).sum())
    if frac == 1.0:
        assert math.isnan(float(loss)) and float(dlogits.float().abs().max()) == 0.0
        return
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel(dlogits, ref_in.grad) < (1e-5 if dtype == torch.float32 else 1e-2)
    ref_px = torch.nn.functional.cross_entropy(logits.float(), target, ignore_index=255, reduction='none')
    assert rel(pixel, ref_px) < 1e-5


def test_cross_entropy_stray_labels_do_not_dilute_the_mean():
    """A label outside [0, C) that is not ignore_index (torch raises on it) carries no loss and no gradient in the
    kernels; it must not be counted in the mean's denominator either."""
    g = gen(5)
    logits = torch.randn(1, 19, 8, 12, generator=g)
    target = torch.randint(0, 19, (1, 8, 12), generator=g)
    target[0, 0, :5] = 255
    stray = target.clone()
    stray[0, 1, :4] = 37
    stray[0, 2, :3] = -2
    clean = stray.clone()
    clean[(stray < 0) | (stray >= 19)] = 255
    loss, dl, _, nvalid = ops.ce_forward(logits.cuda(), stray.cuda(), 255, want_grad=True)
    ref_in = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, clean, ignore_index=255)
    ref.backward()
    assert int(nvalid) == int((clean != 255).sum())
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel(dl, ref_in.grad) < 1e-5


# ------------------------------------------------------------------ confusion matrix ---
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,C,Hi,Wi,frac', [(2, 19, 12, 20, 0.1), (1, 19, 4, 4, 0.0), (2, 19, 3, 5, 0.5), (1, 11, 6, 7, 0.1),
                                             (1, 19, 1, 1, 0.0), (1, 19, 5, 40, 0.1), (2, 19, 24, 24, 1.0)])
def test_fused_head_upsample_cross_entropy(N, C, Hi, Wi, frac, dtype):
    """x8 up-sampling + CE + gradient w.r.t. the low-resolution scores in one kernel, against
    F.interpolate + log_softmax + autograd on the CPU (through the fake backend's formula) and against
    the two-kernel path (tss_upsample_logits_fwd + tss_ce_fwd) on the GPU."""
    g = gen(N * 100 + Hi * 10 + Wi)
    code = _lib.dtype_code(dtype)
    Ho, Wo = 8 * Hi, 8 * Wi
    pitch = 32 if C > 16 else 16
    xc, xg = pair(N, C, Hi, Wi, dtype, g, pitch=pitch, scale=2.0)
    t = torch.randint(0, C, (N, Ho, Wo), generator=g)
    t[torch.rand(N, Ho, Wo, generator=g) < frac] = 255
    lp = (C + 7) // 8 * 8

    def bufs(device):
        return dict(loss_sum=torch.zeros(1, dtype=torch.float64, device=device), nvalid=torch.zeros(1, dtype=torch.int64, device=device),
                    pixel_loss=torch.empty(N, Ho, Wo, device=device), dx32=torch.zeros(N, Hi, Wi, lp, device=device))
    bc, bg = bufs('cpu'), bufs('cuda')
    kw = dict(N=N, C=C, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo, ldx=pitch, ignore_index=255, lddx=lp, ohem=None, dtype=code)
    both('tss_upsample_ce_fwd', dict(x=xc, target=t, **bc, **kw), dict(x=xg, target=t.cuda(), **bg, **kw))
    assert int(bg['nvalid']) == int(bc['nvalid']) == int((t != 255).sum())            # integer: exact
    assert abs(float(bg['loss_sum']) - float(bc['loss_sum'])) <= 1e-5 * max(1.0, abs(float(bc['loss_sum'])))
    assert (bg['pixel_loss'].cpu() - bc['pixel_loss']).abs().max() < 1e-4
    assert rel(bg['dx32'], bc['dx32']) < 1e-4 or float(bc['dx32'].abs().max()) == 0.0
    assert not bg['dx32'][..., C:].any()
    lc, lg = torch.empty(()), torch.empty((), device='cuda')
    dxc, dxg = torch.empty(N, Hi, Wi, lp, dtype=dtype), torch.empty(N, Hi, Wi, lp, dtype=dtype, device='cuda')
    both('tss_upsample_ce_finalize', dict(loss_sum=bc['loss_sum'], nvalid=bc['nvalid'], loss=lc, dx32=bc['dx32'], dx=dxc, n=dxc.numel(), dtype=code),
         dict(loss_sum=bg['loss_sum'], nvalid=bg['nvalid'], loss=lg, dx32=bg['dx32'], dx=dxg, n=dxg.numel(), dtype=code))
    if frac < 1.0:
        assert abs(float(lg) - float(lc)) < 1e-5 * abs(float(lc))
        assert rel(dxg, dxc) < TOL[dtype]
        # same loss as the materialised path (which rounds the logits to the activation dtype first)
        logits = ops.upsample_logits_fwd(xg, Ho, Wo)
        loss2 = ops.ce_forward(logits, t.cuda(), 255, want_grad=False)[0]
        assert abs(float(lg) - float(loss2)) < (1e-5 if dtype == torch.float32 else 5e-3) * abs(float(loss2))
    else:
        assert math.isnan(float(lg)) and math.isnan(float(lc))                          # no valid pixel: NaN like the reference


# ------------------------------------------------------------------ OHEM -----------------
@pytest.mark.parametrize('n,frac,kind', [(24 * 32 * 2, 0.05, 'many_hard'), (24 * 32 * 2, 0.05, 'few_hard'), (100003, 0.1, 'ties'),
                                         (7, 0.5, 'tiny'), (1 << 20, 0.01, 'zeros'), (4099, 0.0, 'empty_top')])
def test_ohem_select_against_sort(n, frac, kind):
    """Radix select + case decision (tss_ohem_select) against the reference's sort-based formula."""
    g = gen(n)
    v = torch.rand(n, generator=g) * 3
    if kind == 'few_hard':
        v = v * 0.05
    if kind == 'ties':
        v = (v * 4).round() / 4                       # heavy ties, also exactly at the selected value
    if kind == 'zeros':
        v[torch.rand(n, generator=g) < 0.995] = 0.0   # (n_keep+1)-th largest is an ignored pixel's 0
    thresh = 0.35667494393873245
    n_keep = int(n * frac)
    srt, _ = torch.sort(v, descending=True)
    want = srt[srt > thresh].mean() if srt[n_keep] > thresh else srt[:n_keep].mean()
    loss, rule = ops.ohem_select(v.cuda(), n_keep, thresh)
    loss2, rule2 = ops.ohem_select(v.cuda(), n_keep, thresh)            # the workspace is left reusable
    torch.cuda.synchronize()
    if n_keep == 0 and not srt[0] > thresh:
        assert math.isnan(float(loss))
    else:
        assert abs(float(loss) - float(want)) <= 1e-6 * abs(float(want)), (float(loss), float(want))
        assert float(loss2) == float(loss) and torch.equal(rule2, rule)
    cut, above, tie, wtie = [float(x) for x in rule.cpu()]
    w = torch.where(v > cut, torch.full_like(v, above), torch.where(v == tie, torch.full_like(v, wtie), torch.zeros_like(v)))
    if not math.isnan(float(loss)):
        assert abs(float((w * v).sum()) - float(want)) <= 1e-4 * abs(float(want))   # the weights reproduce the loss
        assert abs(float(w.sum()) - 1.0) < 1e-4


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('name', ['many_hard', 'few_hard'])
def test_ohem_loss_matches_golden_and_oracle_gradient(name, dtype):
    import os
    from oracle.golden_inputs import ohem_case
    from oracle.init_state import GOLDEN_DIR
    from oracle.losses import ohem as oracle_ohem
    from torch_semantic_segmentation_b200.losses import OHEMLoss
    logits, target, kw = ohem_case(name)
    gold = float(np.load(os.path.join(GOLDEN_DIR, 'ohem.npz'))[name])
    x = logits.to(dtype).cuda().requires_grad_(True)
    loss = OHEMLoss(**kw)(x, target.cuda())
    loss.backward()
    xr = logits.to(dtype).float().requires_grad_(True)
    ref = oracle_ohem(xr, target, **kw)
    ref.backward()
    # per-pixel losses of ~2e-4 come out of logits of magnitude ~12: one fp32 ulp there is ~1e-6, so
    # the bound is 1e-4 relative plus half an ulp of the logits' scale
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref)) + 5e-7
    if dtype == torch.float32:
        assert abs(float(loss) - gold) < 1e-4 * abs(gold) + 5e-7
    # pixels tying with the selected order statistic: the reference's sort keeps an arbitrary subset,
    # the kernel shares the slots evenly (same loss, equally valid subgradient) -> left out
    pl = torch.nn.functional.cross_entropy(logits.to(dtype).float(), target, ignore_index=255, reduction='none')
    srt, _ = torch.sort(pl.flatten(), descending=True)
    vk = srt[int(pl.numel() * kw['numel_frac'])]
    untied = ((pl - vk).abs() > 1e-5 * vk.abs() + 1e-6).unsqueeze(1)     # + the fp32 noise band of the per-pixel losses
    # 'few_hard': softmax - onehot ~ -2e-4 is itself a cancellation of O(1) terms (1e-7 absolute)
    assert rel(x.grad.cpu() * untied, xr.grad * untied) < max(TOL[dtype], 2e-3 if name == 'few_hard' else 0.0)


@pytest.mark.parametrize('n', [0, 1, 7, 4096, 1024 * 2048 + 3])
def test_confusion_from_labels_bit_exact(n):
    rng = np.random.RandomState(n % 1000)
    pred = rng.randint(0, 19, size=n)
    target = rng.randint(0, 19, size=n)
    target[rng.rand(n) < 0.1] = 255
    if n > 8:
        target[3] = -1
        pred[5:64] = 7
        target[5:64] = 7                 # a homogeneous run: exercises the warp aggregation
    cm = torch.zeros(19, 19, dtype=torch.int64).cuda()
    # odd offsets break 16-byte alignment of a view: the wrapper passes a fresh contiguous copy
    ops.confusion_from_labels(torch.from_numpy(pred).cuda(), torch.from_numpy(target).cuda(), 19, cm)
    ops.confusion_from_labels(torch.from_numpy(pred).cuda(), torch.from_numpy(target).cuda(), 19, cm)   # accumulates
    want = o_cm.confusion_matrix(pred, target, 19)
    assert (cm.cpu().numpy() == 2 * want).all()


@pytest.mark.parametrize('dtype', DTYPES)
def test_confusion_from_logits_bit_exact(dtype):
    g = gen(21)
    logits = torch.randn(2, 19, 32, 64, generator=g).to(dtype)
    logits[0, 3, 0, 0] = float('nan')
    logits[0, :, 0, 1] = 1.0                       # all tied -> class 0
    logits[1, 5, 2, 3] = logits[1, 9, 2, 3] = 50.0  # tie -> lowest index
    target = torch.randint(0, 19, (2, 32, 64), generator=g)
    target[torch.rand(2, 32, 64, generator=g) < 0.1] = 255
    cm = torch.zeros(19, 19, dtype=torch.int64).cuda()
    pred = ops.confusion_from_logits(logits.cuda(), target.cuda(), cm, want_pred=True)
    want_pred = logits.float().argmax(1)
    assert torch.equal(pred.cpu(), want_pred)
    assert (o_cm.argmax_classes(logits.float().numpy()) == want_pred.numpy()).all()
    assert (cm.cpu().numpy() == o_cm.confusion_matrix(want_pred.numpy(), target.numpy(), 19)).all()


def test_adamw_flat_arena():
    g = gen(33)
    n = 10007
    p = torch.randn(n, generator=g)
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-3, weight_decay=1e-5)
    pg, m, v = p.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-5, 0, 0, 0], dtype=torch.float32).cuda()
    for step in range(3):
        grad = torch.randn(n, generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        ops.adamw_step(pg, (grad * 2).cuda(), m, v, hyper, grad_scale=0.5)
    assert rel(pg, p_ref) < 1e-6 and float(hyper[5]) == 3.0


# ------------------------------------------------------------------ dense 3x3 as a patch GEMM
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('N,C,H,W', [(2, 128, 6, 9), (1, 16, 1, 1), (3, 8, 5, 2), (1, 128, 32, 64)])
def test_dense3x3_patch_matrix(N, C, H, W, dtype):
    g = gen(N * C + H)
    code = _lib.dtype_code(dtype)
    xc, xg = pair(N, C, H, W, dtype, g)
    cc, cg = pair(N, 9 * C, H, W, dtype, g)
    both('tss_im2col3x3', dict(x=xc, col=cc, N=N, H=H, W=W, C=C, dtype=code), dict(x=xg, col=cg, N=N, H=H, W=W, C=C, dtype=code))
    assert torch.equal(cg.cpu(), cc)                                  # a pure copy: bit-exact
    # the patch GEMM equals the dense convolution
    w = torch.randn(16, C, 3, 3, generator=g)
    wk_c, wk_g = torch.empty(16, 9 * C, 1, 1), torch.empty(16, 9 * C, 1, 1).cuda()
    both('tss_permute_weights3x3', dict(src=w, dst=wk_c, Cout=16, Cin=C, backward=0),
         dict(src=w.cuda(), dst=wk_g, Cout=16, Cin=C, backward=0))
    assert torch.equal(wk_g.cpu(), wk_c)
    ref = torch.nn.functional.conv2d(xc.float(), w, padding=1)
    got = torch.nn.functional.conv2d(cg.cpu().float(), wk_g.cpu())
    assert rel(got, ref) < 1e-5
    # transposes
    dcc, dcg = pair(N, 9 * C, H, W, dtype, g)
    dxc, dxg = pair(N, C, H, W, dtype, g)
    both('tss_col2im3x3', dict(dcol=dcc, dx=dxc, N=N, H=H, W=W, C=C, dtype=code), dict(dcol=dcg, dx=dxg, N=N, H=H, W=W, C=C, dtype=code))
    assert rel(dxg, dxc) < TOL[dtype]
    dwk = torch.randn(16, 9 * C, 1, 1, generator=g)
    dwc = torch.randn(16, C, 3, 3, generator=g)
    dwg = dwc.clone().cuda()
    both('tss_permute_weights3x3', dict(src=dwk, dst=dwc, Cout=16, Cin=C, backward=1),
         dict(src=dwk.cuda(), dst=dwg, Cout=16, Cin=C, backward=1))
    assert torch.equal(dwg.cpu(), dwc)


# ------------------------------------------------------------------ tcgen05 pointwise ---
TC_CASES = [  # K, Nc, M (rows = N*H*W with N=1, H=1)
    (64, 64, 128), (64, 64, 4096), (32, 48, 1000), (48, 64, 777), (64, 384, 2304), (384, 64, 2304), (384, 96, 576),
    (96, 576, 576), (576, 96, 576), (576, 128, 576), (128, 768, 576), (768, 128, 576), (256, 128, 576), (128, 128, 9216),
    (64, 128, 9216), (128, 32, 12), (192, 32, 300), (288, 48, 300), (48, 288, 300), (32, 32, 129), (128, 128, 1),
    (32, 192, 500), (192, 48, 500), (288, 64, 300), (1152, 128, 2048), (32, 64, 700),
    (64, 384, 27648), (64, 384, 25601), (48, 288, 30000), (64, 128, 77000),      # persistent kernel with up to 6 column tiles (K <= 64)
]


@pytest.mark.parametrize('K,Nc,M', TC_CASES)
def test_pwconv_tcgen05(K, Nc, M):
    """impl 1 (TMA + tcgen05.mma + TMEM) against fp32 torch on bf16-rounded operands."""
    g = gen(K + Nc + M)
    dtype, code = torch.bfloat16, 1
    xc, xg = pair(1, K, 1, M, dtype, g)
    w = (torch.randn(Nc, K, 1, 1, generator=g) / math.sqrt(K)).to(dtype).float()     # exactly representable
    wp, wpT = ops.pack_weights_bf16(w.cuda())
    assert torch.equal(wp.float().cpu(), w.view(Nc, K)) and torch.equal(wpT.float().cpu(), w.view(Nc, K).t())
    yc, yg = pair(1, Nc, 1, M, dtype, g)
    sc, sg = torch.zeros(2 * Nc, dtype=torch.float64), torch.zeros(2 * Nc, dtype=torch.float64).cuda()
    base = dict(M=M, K=K, Nc=Nc, ldx=K, ldy=Nc, ldr=0, dtype=code)
    FakeBackend().call('tss_pwconv_fwd', dict(x=xc, w=w, wp=None, y=yc, scale=None, shift=None, res=None, flags=0, stats=sc, impl=0, **base))
    _lib.backend().call('tss_pwconv_fwd', dict(x=xg, w=w.cuda(), wp=wp, y=yg, scale=None, shift=None, res=None, flags=0, stats=sg, impl=1, **base))
    torch.cuda.synchronize()
    assert rel(yg, yc) < 5e-3, rel(yg, yc)          # same operands, fp32 accumulate: only the final rounding differs
    assert rel(sg, sc) < 1e-4
    rc, rg = pair(1, Nc, 1, M, dtype, g)
    scale, shift = torch.rand(Nc, generator=g) + 0.5, torch.randn(Nc, generator=g)
    base['ldr'] = Nc
    FakeBackend().call('tss_pwconv_fwd', dict(x=xc, w=w, wp=None, y=yc, scale=scale, shift=shift, res=rc, flags=1, stats=None, impl=0, **base))
    _lib.backend().call('tss_pwconv_fwd', dict(x=xg, w=w.cuda(), wp=wp, y=yg, scale=scale.cuda(), shift=shift.cuda(), res=rg, flags=1, stats=None, impl=1, **base))
    torch.cuda.synchronize()
    assert rel(yg, yc) < 5e-3
    # dgrad = the same kernel on the transposed pack
    dyc, dyg = pair(1, Nc, 1, M, dtype, g)
    dxc, dxg = pair(1, K, 1, M, dtype, g)
    kw = dict(M=M, K=K, Nc=Nc, lddy=Nc, lddx=K, dtype=code)
    FakeBackend().call('tss_pwconv_dgrad', dict(dy=dyc, w=w, wpT=None, dx=dxc, impl=0, **kw))
    _lib.backend().call('tss_pwconv_dgrad', dict(dy=dyg, w=w.cuda(), wpT=wpT, dx=dxg, impl=1, **kw))
    torch.cuda.synchronize()
    assert rel(dxg, dxc) < 5e-3
    # wgrad on the tensor cores (MN-major operands straight from the activations), accumulate semantics
    dwc = torch.randn(Nc, K, 1, 1, generator=g)
    dwg = dwc.clone().cuda()
    kw = dict(M=M, K=K, Nc=Nc, ldx=K, lddy=Nc, dtype=code, db=None)
    FakeBackend().call('tss_pwconv_wgrad', dict(x=xc, dy=dyc, dw=dwc, impl=0, **kw))
    _lib.backend().call('tss_pwconv_wgrad', dict(x=xg, dy=dyg, dw=dwg, impl=1, **kw))
    torch.cuda.synchronize()
    assert rel(dwg, dwc) < 1e-4, rel(dwg, dwc)
