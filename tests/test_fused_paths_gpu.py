"""The fused / alternative kernel paths behind torch_semantic_segmentation_b200/gates.py, on the GPU.

Round 1 left these built but never run on hardware; round 2 ran all of them on a B200 (gpurun_out/experimental_r2a.log:
73 of 81 passed at the first visit).  Kernel-level tests run by default whether or not the gate they belong to is on
(a gate is on only if it also won the step A/B, gates.py); the few tests still marked ``experimental`` are open items.

* stride-2 depthwise dgrad with the producer's BatchNorm-backward reduction in its epilogue (csrc/dwconv_bnred.cu; FUSE_BNRED_EXT);
* pointwise backward with the BatchNorm-backward apply in the GEMM's A-operand producer (csrc/pwconv_tc_bwd.cu; FUSE_BNAPPLY);
* the same for stride-1 depthwise layers (csrc/dwconv_bwd_fused.cu; FUSE_BNAPPLY_DW);
* the pyramid-pooling branches as grouped launches (csrc/ppm.cu; FUSE_PPM);
* depthwise forward / weight gradient applying the producer's BatchNorm while reading (csrc/dwconv_bnin.cu; FUSE_BNIN);
* pointwise forward applying the producer's BatchNorm in its operand producer (csrc/pwconv_tc_fwd_bnin.cu; FUSE_BNIN_PW);
* BatchNorm finalize folded into the apply kernel (csrc/bn_fused.cu; FUSE_BNFIN);
* the stem convolution and its weight gradient on tcgen05 (csrc/stem_tc.cu; STEM_TC, default on);
* the warp-private confusion-matrix kernel (csrc/metrics.cu; default on);
* the library's own dropout kernel (csrc/dropout.cu; OWN_DROPOUT, default on);
* deferred logits (DEFER_LOGITS, default on: host-side only);
* the device input pipeline (csrc/augment.cu), bit-exact against the OpenCV-pinned CPU pipeline;
* graph replays under a learning-rate schedule, graphed evaluation after graphed training (the benchmarked path)."""
import os

import pytest
import torch

from tests.fake_backend import FakeBackend
from torch_semantic_segmentation_b200 import _lib

pytestmark = pytest.mark.gpu
# still open after the B200 visits of round 2: run with TSS_EXPERIMENTAL=1
experimental = pytest.mark.skipif(os.environ.get('TSS_EXPERIMENTAL') != '1', reason='open item: set TSS_EXPERIMENTAL=1')


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(N, C, H, W, g, dtype):
    base = torch.randn(N, H, W, C, generator=g).to(dtype)
    return base.permute(0, 3, 1, 2), base.cuda().permute(0, 3, 1, 2)


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('C,N,Hi,Wi,relu', [(32, 2, 40, 56, 1), (384, 2, 16, 24, 1), (48, 1, 9, 13, 0), (64, 3, 7, 5, 1), (8, 1, 1, 1, 1)])
def test_dw_dgrad_stride2_with_fused_bn_reduction(C, N, Hi, Wi, relu, dtype):
    g = torch.Generator().manual_seed(C + Hi)
    Ho, Wo = (Hi - 1) // 2 + 1, (Wi - 1) // 2 + 1
    dyc, dyg = nhwc(N, C, Ho, Wo, g, dtype)
    ypc, ypg = nhwc(N, C, Hi, Wi, g, dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    gc, gg = nhwc(N, C, Hi, Wi, g, dtype)
    sc, sg = torch.zeros(2 * C), torch.zeros(2 * C).cuda()
    kw = dict(N=N, Hi=Hi, Wi=Wi, C=C, flags=relu, dtype=_lib.dtype_code(dtype))
    FakeBackend().call('tss_dwconv3x3_dgrad_s2_bnred', dict(dy=dyc, w=w, g=gc, yp=ypc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sc, **kw))
    _lib.backend().call('tss_dwconv3x3_dgrad_s2_bnred', dict(dy=dyg, w=w.cuda(), g=gg, yp=ypg, mean=mean.cuda(), rstd=rstd.cuda(),
                                                            gamma=gamma.cuda(), beta=beta.cuda(), sums=sg, **kw))
    torch.cuda.synchronize()
    tol = 1e-4 if dtype == torch.float32 else 5e-3
    assert rel(gg, gc) < tol, rel(gg, gc)
    assert rel(sg, sc) < 2e-3, rel(sg, sc)


def test_training_step_with_extended_fusion_matches_unfused():
    """Whole training step with the gate on against the gate off.  The bf16 step is not reproducible from run to run to
    better than ~2 % on the gradients in front of a BatchNorm (fp32 atomics in the statistics, then 45 layers of
    amplification: tools/diag_gates.py), so the unfused path's own run-to-run distance is the yardstick."""
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    x, y = train_batch('fastscnn')
    keep = Fn.FUSE_BNRED_EXT

    def run(flag):
        Fn.FUSE_BNRED_EXT = flag
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16).train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            before = _lib.launch_count()
            loss = CrossEntropyLoss(ignore_index=255)(model(x.cuda()), y.cuda())
            loss.backward()
            torch.cuda.synchronize()
            return {k: p.grad.clone() for k, p in model.named_parameters()}, _lib.launch_count() - before, float(loss.detach())
        finally:
            Fn.FUSE_BNRED_EXT = keep
    base, again, fused = run(False), run(False), run(True)
    assert fused[1] == base[1] - 6                  # 4 stride-2 + 2 more stand-alone reductions less
    # the forward pass is untouched; fp64 atomics in the statistics make it reproducible to rounding, not to the bit
    assert abs(fused[2] - base[2]) <= 1e-3 * abs(base[2]) and abs(again[2] - base[2]) <= 1e-3 * abs(base[2])
    for k in ('classifier.3.weight', 'features.0.0.conv1.0.weight', 'downsample.1.0.weight', 'downsample.0.0.weight'):
        noise = rel(again[0][k], base[0][k])
        assert rel(fused[0][k], base[0][k]) < max(3e-2, 3 * noise), (k, noise)


@pytest.mark.parametrize('link', [False, True])
@pytest.mark.parametrize('M_shape,K,Nc,relu', [((2, 16, 24), 64, 384, 1), ((1, 9, 13), 128, 128, 1), ((2, 5, 7), 96, 576, 0),
                                               ((3, 8, 8), 32, 48, 1), ((1, 1, 3), 16, 8, 1), ((2, 32, 64), 48, 64, 1),
                                               ((1, 4, 4), 128, 768, 1)])
def test_pw_backward_with_bn_apply_in_the_operand_producer(M_shape, K, Nc, relu, link):
    """One kernel == tss_bn_bwd_apply -> tss_pwconv_dgrad(_bnred) (emulated on the CPU from the same bf16 inputs)."""
    N, H, W = M_shape
    M = N * H * W
    g = torch.Generator().manual_seed(K + Nc + M)
    dt = torch.bfloat16
    dzc, dzg = nhwc(N, Nc, H, W, g, dt)
    yc, yg = nhwc(N, Nc, H, W, g, dt)
    ypc, ypg = nhwc(N, K, H, W, g, dt)
    par = lambda C: (torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) + 0.5,
                     torch.randn(C, generator=g) * 0.3)
    mean, rstd, gamma, beta = par(Nc)
    pmean, prstd, pgamma, pbeta = par(K)
    wpT = (torch.randn(K, Nc, generator=g) / Nc ** 0.5).to(dt)
    fake = FakeBackend()
    sums = torch.zeros(2 * Nc)
    fake.call('tss_bn_bwd_reduce', dict(dz=dzc, z=None, y=yc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=Nc,
                                        lddz=Nc, ldz=0, ldy=Nc, flags=relu, dtype=1))
    outs = {}
    for name, be, dev in (('cpu', fake, 'cpu'), ('gpu', _lib.backend(), 'cuda')):
        t = lambda v: v.to(dev) if v is not None else None
        dy = torch.zeros(N, H, W, Nc, dtype=dt, device=dev).permute(0, 3, 1, 2)
        dx = torch.zeros(N, H, W, K, dtype=dt, device=dev).permute(0, 3, 1, 2)
        dgamma, dbeta, psums = torch.ones(Nc, device=dev), torch.ones(Nc, device=dev), torch.zeros(2 * K, device=dev)
        be.call('tss_pwconv_bwd_fused', dict(
            dz=dzg if dev == 'cuda' else dzc, y=yg if dev == 'cuda' else yc, lddz=Nc, ldy=Nc, mean=t(mean), rstd=t(rstd),
            gamma=t(gamma), beta=t(beta), sums=t(sums), flags=relu, count=M, dy=dy, lddy=Nc, dgamma=dgamma, dbeta=dbeta,
            wpT=t(wpT), dx=dx, M=M, K=K, Nc=Nc, lddx=K, yp=(ypg if dev == 'cuda' else ypc) if link else None,
            ldyp=K if link else 0, pmean=t(pmean) if link else None, prstd=t(prstd) if link else None,
            pgamma=t(pgamma) if link else None, pbeta=t(pbeta) if link else None, pflags=1 if link else 0,
            psums=psums if link else None))
        outs[name] = (dy, dx, dgamma, dbeta, psums)
    torch.cuda.synchronize()
    c, d = outs['cpu'], outs['gpu']
    assert rel(d[0], c[0]) < 5e-3, ('dy', rel(d[0], c[0]))
    assert rel(d[1], c[1]) < 8e-3, ('dx', rel(d[1], c[1]))
    assert rel(d[2], c[2]) < 1e-6 and rel(d[3], c[3]) < 1e-6
    if link:
        assert rel(d[4], c[4]) < 5e-3, ('psums', rel(d[4], c[4]))


def test_training_step_with_fused_bn_apply_matches_unfused():
    """Whole training step with the gate on against the gate off.  The bf16 step is not reproducible from run to run to
    better than ~2 % on the gradients in front of a BatchNorm (fp32 atomics in the statistics, then 45 layers of
    amplification: tools/diag_gates.py), so the unfused path's own run-to-run distance is the yardstick."""
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    x, y = train_batch('fastscnn')
    keep = Fn.FUSE_BNAPPLY

    def run(flag):
        Fn.FUSE_BNAPPLY = flag
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16).train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            before = _lib.launch_count()
            loss = CrossEntropyLoss(ignore_index=255)(model(x.cuda()), y.cuda())
            loss.backward()
            torch.cuda.synchronize()
            return {k: p.grad.clone() for k, p in model.named_parameters()}, _lib.launch_count() - before, float(loss.detach())
        finally:
            Fn.FUSE_BNAPPLY = keep
    base, again, fused = run(False), run(False), run(True)
    assert fused[1] == base[1] - 22                 # one launch less per residual-free 1x1 layer
    # the forward pass is untouched; fp64 atomics in the statistics make it reproducible to rounding, not to the bit
    assert abs(fused[2] - base[2]) <= 1e-3 * abs(base[2]) and abs(again[2] - base[2]) <= 1e-3 * abs(base[2])
    for k in ('classifier.3.weight', 'features.0.0.conv1.0.weight', 'downsample.1.0.weight', 'downsample.0.0.weight'):
        noise = rel(again[0][k], base[0][k])
        assert rel(fused[0][k], base[0][k]) < max(3e-2, 3 * noise), (k, noise)


def test_device_input_pipeline_is_bit_exact_with_the_cpu_pipeline():
    import numpy as np
    from oracle import augment as A
    from torch_semantic_segmentation_b200.data import DeviceTransform, eval_transform
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'augment.npz'))
    for seed, h, w, scale, hf, wf, flip, crop in A.GOLDEN_CASES:
        img, lab = A.sample(seed, h, w)
        t = DeviceTransform(crop=crop)
        x, y = t(torch.from_numpy(img)[None].cuda(), torch.from_numpy(lab)[None].cuda(), draws=[(scale, hf, wf, bool(flip))])
        np.testing.assert_array_equal(x[0].cpu().numpy(), gold['image_%d' % seed])
        np.testing.assert_array_equal(y[0].cpu().numpy(), gold['label_%d' % seed].astype(np.int64))
    # a Cityscapes-sized batch: the reference's crop at both ends of its scale range, and the evaluation transform
    samples = [A.sample(40 + i, 1024, 2048) for i in range(2)]
    images = torch.from_numpy(np.stack([s[0] for s in samples])).cuda()
    labels = torch.from_numpy(np.stack([s[1] for s in samples])).cuda()
    draws = [(1.5, 0.25, 0.75, True), (3.0, 0.9, 0.1, False)]
    x, y = DeviceTransform(crop=(512, 768))(images, labels, draws=draws)
    for i, d in enumerate(draws):
        ex, ey = A.train_transform(samples[i][0], samples[i][1], d[0], d[1], d[2], d[3], (512, 768))
        np.testing.assert_array_equal(x[i].cpu().numpy(), ex)
        np.testing.assert_array_equal(y[i].cpu().numpy(), ey)
    x, y = eval_transform()(images, labels)
    ex, ey = A.eval_transform(*samples[1])
    np.testing.assert_array_equal(x[1].cpu().numpy(), ex)
    np.testing.assert_array_equal(y[1].cpu().numpy(), ey)


@experimental
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(12, 24, 24), (2, 32, 64), (3, 5, 7), (2, 1, 1)])
def test_grouped_pyramid_pooling_matches_layer_by_layer(shape, dtype):
    from torch_semantic_segmentation_b200 import functional as Fn, ops
    from torch_semantic_segmentation_b200.models.fastscnn import PyramidPoolingModule
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    N, H, W = shape
    keep = Fn.FUSE_PPM
    runs = {}
    try:
        for flag in (False, True):
            Fn.FUSE_PPM = flag
            torch.manual_seed(0)
            m = set_compute_dtype(PyramidPoolingModule(128, 128), dtype, pw_impl=0).cuda().train()
            FlatAdamW(m.parameters(), lr=1e-3).zero_grad()
            g = torch.Generator().manual_seed(1)
            x = ops.as_nhwc(torch.randn(N, 128, H, W, generator=g).to(dtype).cuda()).requires_grad_()
            before = _lib.launch_count()
            out = m(x)
            (out.float() * torch.randn(out.shape, generator=g).cuda()).sum().backward()
            torch.cuda.synchronize()
            runs[flag] = (out.detach(), x.grad, {k: p.grad.clone() for k, p in m.named_parameters()},
                          {k: v.clone().float() for k, v in m.state_dict().items() if 'running' in k or 'tracked' in k},
                          _lib.launch_count() - before)
    finally:
        Fn.FUSE_PPM = keep
    a, b = runs[False], runs[True]
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert rel(b[0], a[0]) < tol and rel(b[1], a[1]) < 2 * tol
    for k in a[2]:
        assert rel(b[2][k], a[2][k]) < 3 * tol, k
    for k in a[3]:
        assert rel(b[3][k], a[3][k]) < tol, k
    assert b[4] <= a[4] - 30


def test_training_step_with_grouped_pyramid_pooling_matches_default():
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    x, y = train_batch('fastscnn')
    out = {}
    keep = Fn.FUSE_PPM
    for flag in (False, True):
        Fn.FUSE_PPM = flag
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            FlatAdamW(model.parameters(), lr=1e-3).zero_grad()
            logits = model(x.cuda())
            loss = CrossEntropyLoss(ignore_index=255)(logits, y.cuda())
            loss.backward()
            torch.cuda.synchronize()
            out[flag] = (float(loss), logits.detach(), model.classifier[3].weight.grad.clone())
        finally:
            Fn.FUSE_PPM = keep
    assert abs(out[True][0] - out[False][0]) < 1e-4 * abs(out[False][0])
    assert rel(out[True][1], out[False][1]) < 1e-4 and rel(out[True][2], out[False][2]) < 1e-3


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_training_step_with_one_launch_batchnorm_backward_matches_default(dtype):
    """BN_BWD_ONEPASS: every stand-alone reduce + apply pair of the step becomes one launch with a grid barrier; same loss
    and gradients (fp32: to rounding of the atomics' order), fewer launches, and no barrier ever timed out."""
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import ops
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    x, y = train_batch('fastscnn')
    out = {}
    keep = ops.BN_BWD_ONEPASS
    for flag in (False, True):
        ops.BN_BWD_ONEPASS = flag
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().set_compute_dtype(dtype).train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            before = _lib.launch_count()
            for _ in range(2):
                model.zero_grad()
                logits = model(x.cuda())
                loss = CrossEntropyLoss(ignore_index=255)(logits, y.cuda())
                loss.backward()
            torch.cuda.synchronize()
            out[flag] = (float(loss), logits.detach().float(), model.classifier[3].weight.grad.clone(),
                         model.downsample[0][0].weight.grad.clone(), _lib.launch_count() - before)
        finally:
            ops.BN_BWD_ONEPASS = keep
    assert not ops.grid_sync_timed_out()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert abs(out[True][0] - out[False][0]) <= tol * abs(out[False][0])
    assert rel(out[True][1], out[False][1]) < tol and rel(out[True][2], out[False][2]) < 10 * tol
    if dtype == torch.float32:
        assert rel(out[True][3], out[False][3]) < 5e-2          # the stem's gradient: 45 layers of amplified atomics noise
    assert out[True][4] <= out[False][4] - 2 * 10, (out[True][4], out[False][4])


def test_training_step_with_finalize_folded_into_apply_matches_default():
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    x, y = train_batch('fastscnn')
    out = {}
    keep = Fn.FUSE_BNFIN, Fn.FUSE_BNIN
    for flag in (False, True):
        Fn.FUSE_BNFIN, Fn.FUSE_BNIN = flag, False               # (the launch count below: no apply pass handed to a consumer)
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            before = _lib.launch_count()
            for _ in range(2):                                  # twice: the scratch / ticket must be clean for the second step
                model.zero_grad()
                logits = model(x.cuda())
                loss = CrossEntropyLoss(ignore_index=255)(logits, y.cuda())
                loss.backward()
            torch.cuda.synchronize()
            out[flag] = (float(loss), logits.detach(), model.classifier[3].weight.grad.clone(),
                         model.downsample[0][1].running_var.clone(), _lib.launch_count() - before)
        finally:
            Fn.FUSE_BNFIN, Fn.FUSE_BNIN = keep
    # fp32 mode: both paths do the same arithmetic (fp64 statistics -> fp32 scale / shift), equal to fp32 rounding
    assert abs(out[True][0] - out[False][0]) < 1e-4 * abs(out[False][0])
    assert rel(out[True][1], out[False][1]) < 1e-4 and rel(out[True][2], out[False][2]) < 1e-3
    assert rel(out[True][3], out[False][3]) < 1e-6
    assert out[True][4] == out[False][4] - 2 * 44


@pytest.mark.parametrize('N,H,W', [(2, 64, 96), (1, 32, 300), (1, 7, 9), (3, 33, 515), (12, 768, 768)])
def test_stem_weight_gradient_through_the_patch_matrix(N, H, W):
    """tss_stem3x3s2_wgrad_patches (coalesced patch matrix + the TMA-fed pointwise weight-gradient GEMM) against the SIMT
    stem weight gradient; the patch matrix itself against F.unfold."""
    g = torch.Generator().manual_seed(H * 3 + W)
    x = torch.randn(N, 3, H, W, generator=g).cuda()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dy = torch.randn(N, Ho, Wo, 32, generator=g).to(torch.bfloat16).cuda().permute(0, 3, 1, 2)
    be = _lib.backend()
    ref = torch.zeros(32, 3, 3, 3, device='cuda')
    be.call('tss_stem3x3s2_wgrad', dict(x=x, dy=dy, dw=ref, N=N, H=H, W=W, Cout=32, dtype=1))
    dw = torch.full((32, 3, 3, 3), 0.5, device='cuda')                       # accumulates into what is there
    patches = torch.full((N * Ho * Wo, 32), float('nan'), dtype=torch.bfloat16, device='cuda')
    dw32 = torch.full((32, 32), float('nan'), device='cuda')
    be.call('tss_stem3x3s2_wgrad_patches', dict(x=x, dy=dy, patches=patches, dw32=dw32, dw=dw, N=N, H=H, W=W, Cout=32))
    torch.cuda.synchronize()
    want = torch.nn.functional.unfold(x, 3, padding=1, stride=2).transpose(1, 2).reshape(-1, 27)      # (ci, ky, kx) order
    assert torch.equal(patches[:, :27].float(), want.to(torch.bfloat16).float()) and float(patches[:, 27:].abs().max()) == 0.0
    assert rel(dw - 0.5, ref) < 1e-2                       # the image is rounded to bf16 on this path


@pytest.mark.parametrize('N,H,W', [(2, 64, 96), (1, 32, 300), (1, 7, 9), (12, 768, 768)])
def test_stem_on_tensor_cores_matches_the_simt_stem(N, H, W):
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(N, 3, H, W, generator=g).cuda()
    w = (torch.randn(32, 3, 3, 3, generator=g) / 5).cuda()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    be = _lib.backend()
    ys, sts, dws = [], [], []
    dy = torch.randn(N, Ho, Wo, 32, generator=g).to(torch.bfloat16).cuda().permute(0, 3, 1, 2)
    for tc in (False, True):
        y = torch.zeros(N, Ho, Wo, 32, dtype=torch.bfloat16, device='cuda').permute(0, 3, 1, 2)
        stats = torch.zeros(64, dtype=torch.float64, device='cuda')
        dw = torch.zeros(32, 3, 3, 3, device='cuda')
        if tc:
            be.call('tss_stem3x3s2_fwd_tc', dict(x=x, w=w, y=y, N=N, H=H, W=W, Cout=32, scale=None, shift=None, flags=0, stats=stats))
            be.call('tss_stem3x3s2_wgrad_tc', dict(x=x, dy=dy, dw=dw, N=N, H=H, W=W, Cout=32))
        else:
            be.call('tss_stem3x3s2_fwd', dict(x=x, w=w, y=y, N=N, H=H, W=W, Cout=32, scale=None, shift=None, flags=0, stats=stats, dtype=1))
            be.call('tss_stem3x3s2_wgrad', dict(x=x, dy=dy, dw=dw, N=N, H=H, W=W, Cout=32, dtype=1))
        torch.cuda.synchronize()
        ys.append(y.float()); sts.append(stats); dws.append(dw)
    assert rel(ys[1], ys[0]) < 1e-2                       # image and weights rounded to bf16 on the tensor-core path
    assert rel(sts[1], sts[0]) < 1e-2
    assert rel(dws[1], dws[0]) < 1e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_grouped_pyramid_pooling_eval_matches_layer_by_layer(dtype):
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models import fastscnn
    keep = (Fn.FUSE_PPM, Fn.FUSE_PPM_EVAL)
    outs = {}
    try:
        for flag in (False, True):
            Fn.FUSE_PPM = Fn.FUSE_PPM_EVAL = flag
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().set_compute_dtype(dtype).eval()
            g = torch.Generator().manual_seed(3)
            x = torch.randn(1, 3, 256, 512, generator=g).cuda()
            before = _lib.launch_count()
            with torch.no_grad():
                outs[flag] = (model(x).float(), _lib.launch_count() - before)
            torch.cuda.synchronize()
    finally:
        Fn.FUSE_PPM, Fn.FUSE_PPM_EVAL = keep
    assert rel(outs[True][0], outs[False][0]) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert outs[True][1] <= outs[False][1] - 7


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('C,N,H,W,relu', [(384, 2, 48, 48, 1), (64, 2, 12, 20, 1), (96, 1, 9, 40, 0), (128, 1, 17, 33, 1), (576, 2, 24, 24, 1)])
def test_dw_backward_with_bn_apply_matches_the_two_kernel_path(C, N, H, W, relu, dtype):
    """tss_dwconv3x3_bwd_fused == tss_bn_bwd_apply -> tss_dwconv3x3_dgrad_bnred, both on the GPU."""
    g = torch.Generator().manual_seed(C + H + W)
    mk = lambda: torch.randn(N, H, W, C, generator=g).to(dtype).cuda().permute(0, 3, 1, 2)
    dz, y, yp = mk(), mk(), mk()
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).cuda()
    par = lambda: [t.cuda() for t in (torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5,
                                      torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3)]
    mean, rstd, gamma, beta = par()
    pmean, prstd, pgamma, pbeta = par()
    M = N * H * W
    be = _lib.backend()
    code = _lib.dtype_code(dtype)
    sums = torch.zeros(2 * C, device='cuda')
    be.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=C, lddz=C,
                                      ldz=0, ldy=C, flags=relu, dtype=code))
    new = lambda: torch.zeros(N, H, W, C, dtype=dtype, device='cuda').permute(0, 3, 1, 2)
    dy1, g1, dy2, g2 = new(), new(), new(), new()
    dga1, dbe1, ps1 = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda'), torch.zeros(2 * C, device='cuda')
    dga2, dbe2, ps2 = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda'), torch.zeros(2 * C, device='cuda')
    be.call('tss_bn_bwd_apply', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, dy=dy1, dres=None,
                                     dgamma=dga1, dbeta=dbe1, M=M, count=M, C=C, lddz=C, ldz=0, ldy=C, lddy=C, lddres=0, flags=relu,
                                     dtype=code))
    be.call('tss_dwconv3x3_dgrad_bnred', dict(dy=dy1, w=w, g=g1, N=N, H=H, W=W, C=C, yp=yp, mean=pmean, rstd=prstd, gamma=pgamma,
                                              beta=pbeta, flags=1, sums=ps1, dtype=code))
    be.call('tss_dwconv3x3_bwd_fused', dict(dz=dz, y=y, w=w, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu,
                                            count=M, dy=dy2, dgamma=dga2, dbeta=dbe2, g=g2, N=N, H=H, W=W, C=C, yp=yp, pmean=pmean,
                                            prstd=prstd, pgamma=pgamma, pbeta=pbeta, pflags=1, psums=ps2, dtype=code))
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel(dy2, dy1) < tol and rel(g2, g1) < 2 * tol
    assert rel(dga2, dga1) < 1e-6 and rel(dbe2, dbe1) < 1e-6 and rel(ps2, ps1) < max(tol, 2e-3)


@pytest.mark.parametrize('kind', ['random', 'piecewise'])
def test_warp_private_confusion_matrix_is_exact(kind, monkeypatch):
    import numpy as np
    from oracle import confusion as o_cm
    monkeypatch.setenv('TSS_CM_VARIANT', '1')
    g = torch.Generator().manual_seed(11)
    n = 1024 * 2048 + 3
    if kind == 'piecewise':
        pred = torch.randint(0, 19, (n // 97 + 1,), generator=g).repeat_interleave(97)[:n]
        label = torch.randint(0, 19, (n // 211 + 1,), generator=g).repeat_interleave(211)[:n]
    else:
        pred, label = torch.randint(0, 19, (n,), generator=g), torch.randint(0, 19, (n,), generator=g)
    label = label.clone()
    label[torch.rand(n, generator=g) < 0.1] = 255
    cm = torch.zeros(19, 19, dtype=torch.int64, device='cuda')
    for _ in range(3):                                     # accumulates
        _lib.backend().call('tss_confusion_from_labels', dict(pred=pred.cuda(), target=label.cuda(), n=n, C=19, cm=cm))
    torch.cuda.synchronize()
    want = 3 * o_cm.confusion_matrix(pred.numpy(), label.numpy(), 19)
    assert np.array_equal(cm.cpu().numpy(), want)


def test_deferred_logits_training_step_matches_default():
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    x, y = train_batch('fastscnn')
    loss_fn = CrossEntropyLoss(ignore_index=255)
    keep = Fn.DEFER_LOGITS
    out = {}
    try:
        for flag in (False, True):
            Fn.DEFER_LOGITS = flag
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16).train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            assert Fn.enable_deferred_logits(model, loss_fn) == flag
            before = _lib.launch_count()
            loss = loss_fn(model(x.cuda()), y.cuda())
            loss.backward()
            torch.cuda.synchronize()
            out[flag] = (float(loss), model.classifier[3].weight.grad.clone(), _lib.launch_count() - before)
    finally:
        Fn.DEFER_LOGITS = keep
    # same kernels, same inputs: equal up to the order of the fp64 atomics in the BatchNorm statistics
    assert abs(out[True][0] - out[False][0]) <= 1e-3 * abs(out[False][0]) and rel(out[True][1], out[False][1]) < 2e-2
    assert out[True][2] == out[False][2] - 1


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_own_dropout_matches_the_philox_reference(dtype):
    from tests.philox_ref import keep_mask
    from torch_semantic_segmentation_b200 import ops
    N, C, H, W, p = 3, 128, 96, 96, 0.1
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, H, W, C, generator=g).to(dtype).cuda().permute(0, 3, 1, 2)
    rng = ops.rng_state('cuda', seed=987654321)
    for step in range(2):
        y, used = ops.dropout_fwd(x, p)
        dx = ops.dropout_bwd(x, p, used)
        torch.cuda.synchronize()
        assert int(used) == step and int(rng[1]) == step + 1 and int(rng[2]) == 0
        keep, scale = keep_mask(987654321, step, x.numel(), p)
        keep = torch.from_numpy(keep).view(N, H, W, C).permute(0, 3, 1, 2).cuda()
        want = torch.where(keep, x.float() * float(scale), torch.zeros((), device='cuda')).to(dtype)
        assert torch.equal(y, want) and torch.equal(dx, want)


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('C,N,H,W,stride', [(384, 2, 96, 96, 2), (384, 2, 48, 48, 1), (576, 2, 24, 24, 1), (96, 1, 9, 40, 1), (32, 2, 16, 24, 2)])
def test_dw_kernels_with_input_batchnorm_match_apply_then_conv(C, N, H, W, stride, dtype):
    """tss_dwconv3x3_fwd_bnin / wgrad_bnin == tss_bn_apply followed by tss_dwconv3x3_fwd / wgrad, both on the GPU."""
    g = torch.Generator().manual_seed(C + H + stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    mk = lambda h, w_: torch.randn(N, h, w_, C, generator=g).to(dtype).cuda().permute(0, 3, 1, 2)
    x, dy = mk(H, W), mk(Ho, Wo)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).cuda()
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.5 + 0.3).cuda()
    be = _lib.backend()
    code = _lib.dtype_code(dtype)
    new = lambda h, w_: torch.zeros(N, h, w_, C, dtype=dtype, device='cuda').permute(0, 3, 1, 2)
    z, y1, y2 = new(H, W), new(Ho, Wo), new(Ho, Wo)
    st1, st2 = torch.zeros(2 * C, dtype=torch.float64, device='cuda'), torch.zeros(2 * C, dtype=torch.float64, device='cuda')
    dw1, dw2 = torch.zeros(C, 1, 3, 3, device='cuda'), torch.zeros(C, 1, 3, 3, device='cuda')
    M = N * H * W
    be.call('tss_bn_apply', dict(y=x, scale=sc, shift=sh, y2=None, scale2=None, shift2=None, res=None, z=z, M=M, C=C, ldy=C, ldy2=0,
                                 ldr=0, ldz=C, flags=1, dtype=code))
    be.call('tss_dwconv3x3_fwd', dict(x=z, w=w, y=y1, N=N, Hi=H, Wi=W, C=C, stride=stride, dilation=1, scale=None, shift=None, flags=0,
                                      stats=st1, dtype=code))
    be.call('tss_dwconv3x3_wgrad', dict(x=z, dy=dy, dw=dw1, N=N, Hi=H, Wi=W, C=C, stride=stride, dilation=1, dtype=code))
    be.call('tss_dwconv3x3_fwd_bnin', dict(x=x, in_scale=sc, in_shift=sh, in_flags=1, w=w, y=y2, N=N, Hi=H, Wi=W, C=C, stride=stride,
                                           stats=st2, dtype=code))
    be.call('tss_dwconv3x3_wgrad_bnin', dict(x=x, in_scale=sc, in_shift=sh, in_flags=1, dy=dy, dw=dw2, N=N, Hi=H, Wi=W, C=C,
                                             stride=stride, dtype=code))
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 8e-3           # bf16: the unfused path rounds z to bf16 before the stencil
    assert rel(y2, y1) < tol and rel(st2, st1) < tol and rel(dw2, dw1) < tol


@pytest.mark.parametrize('M_shape,K,Nc,relu', [((12, 48, 48), 384, 64, 1), ((12, 24, 24), 576, 96, 1), ((12, 24, 24), 768, 128, 1),
                                               ((2, 5, 7), 576, 96, 0), ((1, 1, 3), 16, 16, 1)])
def test_pw_forward_with_input_batchnorm_matches_apply_then_gemm(M_shape, K, Nc, relu):
    """tss_pwconv_fwd_bnin == tss_bn_apply followed by tss_pwconv_fwd (impl 1), both on the GPU."""
    N, H, W = M_shape
    M = N * H * W
    g = torch.Generator().manual_seed(K + Nc + M)
    dt = torch.bfloat16
    x = torch.randn(N, H, W, K, generator=g).to(dt).cuda().permute(0, 3, 1, 2)
    sc, sh = (torch.rand(K, generator=g) + 0.5).cuda(), (torch.randn(K, generator=g) * 0.4).cuda()
    w = (torch.randn(Nc, K, 1, 1, generator=g) / K ** 0.5).cuda()
    wp = w.view(Nc, K).to(dt).contiguous()
    be = _lib.backend()
    new = lambda C: torch.zeros(N, H, W, C, dtype=dt, device='cuda').permute(0, 3, 1, 2)
    z1, y1, z2, y2 = new(K), new(Nc), new(K), new(Nc)
    st1, st2 = torch.zeros(2 * Nc, dtype=torch.float64, device='cuda'), torch.zeros(2 * Nc, dtype=torch.float64, device='cuda')
    be.call('tss_bn_apply', dict(y=x, scale=sc, shift=sh, y2=None, scale2=None, shift2=None, res=None, z=z1, M=M, C=K, ldy=K, ldy2=0,
                                 ldr=0, ldz=K, flags=relu, dtype=1))
    be.call('tss_pwconv_fwd', dict(x=z1, w=w, wp=wp, y=y1, M=M, K=K, Nc=Nc, ldx=K, ldy=Nc, scale=None, shift=None, res=None, ldr=0,
                                   flags=0, stats=st1, impl=1, dtype=1))
    be.call('tss_pwconv_fwd_bnin', dict(x=x, ldx=K, in_scale=sc, in_shift=sh, in_flags=relu, z=z2, ldz=K, wp=wp, y=y2, ldy=Nc, M=M, K=K,
                                        Nc=Nc, stats=st2))
    torch.cuda.synchronize()
    assert rel(z2, z1) < 1e-6 and rel(y2, y1) < 5e-3 and rel(st2, st1) < 1e-3


def test_graph_replays_follow_a_learning_rate_schedule():
    """Default path (FlatAdamW + cuda_graph=True): param_groups['lr'] changed between replays reaches the device
    (GraphedTrainStep refreshes the pinned staging buffer before each replay).  Parked here until it has run once."""
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200.engine import GraphedTrainStep
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    x, y = train_batch('fastscnn')
    x, y = x.cuda(), y.cuda()
    torch.manual_seed(0)
    model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16).train()
    opt = FlatAdamW(model.parameters(), lr=1e-3)
    step = GraphedTrainStep(model, opt, CrossEntropyLoss(ignore_index=255), x, y)
    step(x, y)
    torch.cuda.synchronize()
    assert abs(float(opt.hyper[0]) - 1e-3) < 1e-9
    opt.param_groups[0]['lr'] = 0.0
    before = opt.param_arena.clone()
    step(x, y)
    torch.cuda.synchronize()
    assert float(opt.hyper[0]) == 0.0
    assert torch.equal(opt.param_arena, before)              # lr 0 and decoupled weight decay lr*wd = 0: nothing moves


def test_one_graph_per_staging_slot_trains_like_the_copying_path():
    """engine.SLOT_GRAPHS (host side only): graphs bound to the two staging slots read their batch in place; six different
    batches through the trainer end with the same parameters as the copying path -- up to the run-to-run noise of six
    bf16 optimisation steps, which the copying path's own repeat measures -- with eager and with lazy loss read-back."""
    import torch_semantic_segmentation_b200.engine as E
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(2, 3, 64, 96, generator=g).pin_memory(), torch.randint(0, 19, (2, 64, 96), generator=g).pin_memory())
               for _ in range(6)]
    keep = E.SLOT_GRAPHS

    def run(flag, lazy=False):
        E.SLOT_GRAPHS = flag
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            opt = FlatAdamW(model.parameters(), lr=1e-3)
            trainer = E.create_segmentation_trainer(model, opt, CrossEntropyLoss(ignore_index=255), 'cuda', use_f16=True,
                                                    logging=False, cuda_graph=True, lazy_loss=lazy)
            outputs = []
            trainer.add_event_handler(E.Events.ITERATION_COMPLETED, lambda e: outputs.append(e.state.output))
            state = trainer.run(batches, max_epochs=1)
            torch.cuda.synchronize()
            return outputs, opt.param_arena.clone(), state.iteration
        finally:
            E.SLOT_GRAPHS = keep
    base, again, slots, lazy = run(False), run(False), run(True), run(True, lazy=True)
    assert base[2] == slots[2] == lazy[2] == 6
    noise = rel(again[1], base[1])
    # the first batch gets 3 warm-up steps + capture in both modes; the slot-1 graph is captured without warm-up steps
    assert rel(slots[1], base[1]) < max(5e-3, 3 * noise), (rel(slots[1], base[1]), noise)
    assert rel(lazy[1], base[1]) < max(5e-3, 3 * noise)
    assert abs(slots[0][-1] - base[0][-1]) < 5e-2 * abs(base[0][-1])
    # lazy read-back: iteration i reports the loss of iteration i-1 (the first one its own)
    assert all(abs(a - b) < 5e-2 * abs(b) for a, b in zip(lazy[0][2:], slots[0][1:-1]))


@pytest.mark.parametrize('N,H,W', [(2, 64, 96), (1, 32, 300), (12, 768, 768)])
def test_stem_wgrad_with_bn_apply_matches_apply_then_wgrad(N, H, W):
    """tss_stem3x3s2_wgrad_tc_bn == tss_bn_bwd_apply followed by tss_stem3x3s2_wgrad_tc, both on the GPU."""
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(N, 3, H, W, generator=g).cuda()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    mk = lambda: torch.randn(N, Ho, Wo, 32, generator=g).to(torch.bfloat16).cuda().permute(0, 3, 1, 2)
    dz, y = mk(), mk()
    mean, rstd = (torch.randn(32, generator=g) * 0.2).cuda(), (torch.rand(32, generator=g) + 0.5).cuda()
    gamma, beta = (torch.rand(32, generator=g) + 0.5).cuda(), (torch.randn(32, generator=g) * 0.3).cuda()
    M = N * Ho * Wo
    be = _lib.backend()
    sums = torch.zeros(64, device='cuda')
    be.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=32, lddz=32,
                                      ldz=0, ldy=32, flags=1, dtype=1))
    dy = torch.zeros(N, Ho, Wo, 32, dtype=torch.bfloat16, device='cuda').permute(0, 3, 1, 2)
    dw1, dw2 = torch.zeros(32, 3, 3, 3, device='cuda'), torch.zeros(32, 3, 3, 3, device='cuda')
    dg1, db1, dg2, db2 = (torch.zeros(32, device='cuda') for _ in range(4))
    be.call('tss_bn_bwd_apply', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, dy=dy, dres=None,
                                     dgamma=dg1, dbeta=db1, M=M, count=M, C=32, lddz=32, ldz=0, ldy=32, lddy=32, lddres=0, flags=1, dtype=1))
    be.call('tss_stem3x3s2_wgrad_tc', dict(x=x, dy=dy, dw=dw1, N=N, H=H, W=W, Cout=32))
    be.call('tss_stem3x3s2_wgrad_tc_bn', dict(x=x, dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=1, count=M,
                                              dw=dw2, dgamma=dg2, dbeta=db2, N=N, H=H, W=W, Cout=32))
    torch.cuda.synchronize()
    assert rel(dw2, dw1) < 5e-3 and rel(dg2, dg1) < 1e-6 and rel(db2, db1) < 1e-6


@pytest.mark.parametrize('arch', ['fastscnn', 'contextnet14'])
def test_all_training_gates_together_on_the_gpu(arch):
    """Every off-by-default training gate at once, bf16: three optimisation steps through the real kernels agree with the
    default path at bf16 level, with far fewer launches (the GPU counterpart of
    tests/test_host_logic.py::test_all_gates_together_train_and_eval)."""
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    flags = ['FUSE_BNRED_EXT', 'FUSE_BNAPPLY', 'FUSE_BNAPPLY_DW', 'FUSE_PPM', 'FUSE_BNFIN', 'FUSE_BNIN', 'FUSE_BNIN_PW', 'STEM_TC',
             'STEM_BWD_FUSED', 'DEFER_LOGITS', 'OWN_DROPOUT']
    keep = {f: getattr(Fn, f) for f in flags}
    factory = {'fastscnn': fastscnn, 'contextnet14': contextnet14}[arch]
    x, y = train_batch('fastscnn')
    x, y = x.cuda(), y.cuda()
    runs = {}
    try:
        for on in (False, True):
            for f in flags:
                setattr(Fn, f, on)
            torch.manual_seed(0)
            model = factory(3, 19).cuda().set_compute_dtype(torch.bfloat16).train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            opt = FlatAdamW(model.parameters(), lr=1e-3)
            loss_fn = CrossEntropyLoss(ignore_index=255)
            Fn.enable_deferred_logits(model, loss_fn)
            before, losses = _lib.launch_count(), []
            for _ in range(3):
                opt.zero_grad()
                loss = loss_fn(model(x), y)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
            torch.cuda.synchronize()
            runs[on] = (losses, _lib.launch_count() - before, bool(torch.isfinite(opt.param_arena).all()))
    finally:
        for f, v in keep.items():
            setattr(Fn, f, v)
    assert runs[True][2] and runs[False][2]
    for a, b in zip(runs[True][0], runs[False][0]):
        assert abs(a - b) < 1e-2 * abs(b), (runs[True][0], runs[False][0])
    assert runs[True][1] < runs[False][1] - 200          # three steps


def test_graphed_eval_follows_graphed_training():
    """Default path: train with CUDA-graph replays, evaluate with the graphed evaluator, train on, evaluate again -- the
    second evaluation must see the new weights and BatchNorm statistics (eval operands refreshed in place;
    GraphedTrainStep bumps ops.WEIGHTS_EPOCH).  Checked against an eager evaluation of the same model."""
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200.engine import GraphedTrainStep, create_segmentation_evaluator
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    x, y = train_batch('fastscnn')
    xd, yd = x.cuda(), y.cuda()
    torch.manual_seed(0)
    model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16)
    opt = FlatAdamW(model.parameters(), lr=5e-3)
    step = GraphedTrainStep(model.train(), opt, CrossEntropyLoss(ignore_index=255), xd, yd)
    graphed = create_segmentation_evaluator(model, 'cuda', num_classes=19, cuda_graph=True)
    eager = create_segmentation_evaluator(model, 'cuda', num_classes=19, cuda_graph=False)
    cms = []
    for _round in range(2):
        model.train()
        for _ in range(5):
            step(xd, yd)
        torch.cuda.synchronize()
        a = graphed.run([(x, y)]).metrics['confusion_matrix'].clone()
        b = eager.run([(x, y)]).metrics['confusion_matrix'].clone()
        assert torch.equal(a, b), _round
        cms.append(a)
    assert not torch.equal(cms[0], cms[1])                     # training moved the predictions between the two evaluations


@pytest.mark.parametrize('Nc,shape', [(19, (12, 96, 96)), (19, (2, 12, 20)), (11, (1, 5, 7)), (21, (3, 9, 13))])
def test_class_scores_on_tensor_cores_match_the_simt_gemms(Nc, shape):
    """nn.Conv2d(128, Nc, 1) with bias (fastscnn.py:97) through the tcgen05 GEMMs with zero-padded operands
    (functional.CLASS_TC) against the SIMT fp32-accumulate GEMMs: forward, dgrad, wgrad, bias gradient."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200 import ops
    N, H, W = shape
    K = 128
    g = torch.Generator().manual_seed(Nc + H)
    x = torch.randn(N, H, W, K, generator=g).to(torch.bfloat16).cuda().permute(0, 3, 1, 2)
    w = (torch.randn(Nc, K, 1, 1, generator=g) / K ** 0.5).cuda()
    b = torch.randn(Nc, generator=g).cuda()
    pitch = (Nc + 7) // 8 * 8                                    # what the fused head hands back
    dy_buf = torch.zeros(N, H, W, pitch, dtype=torch.bfloat16, device='cuda')
    dy_buf[..., :Nc] = torch.randn(N, H, W, Nc, generator=g).to(torch.bfloat16).cuda()
    dy = dy_buf[..., :Nc].permute(0, 3, 1, 2)
    keep = Fn.CLASS_TC
    out = {}
    try:
        for flag in (False, True):
            Fn.CLASS_TC = flag
            xi = x.clone().requires_grad_(True)
            wi, bi = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            before = _lib.launch_count()
            y = Fn.ConvBias.apply(xi, wi, bi)
            y.backward(dy)
            torch.cuda.synchronize()
            out[flag] = (y.detach().float().clone(), xi.grad.float(), wi.grad.clone(), bi.grad.clone(), _lib.launch_count() - before)
    finally:
        Fn.CLASS_TC = keep
    ref = torch.nn.functional.conv2d(x.float(), w, b)
    assert rel(out[True][0], ref) < 5e-3 and rel(out[False][0], ref) < 5e-3
    assert ops.geom(Fn.ConvBias.apply(x, w, b))[4] >= 32 or Nc <= 16
    for i, tol in ((1, 8e-3), (2, 1e-3), (3, 1e-4)):              # dx is rounded to bf16; dw, db are fp32 accumulations
        assert rel(out[True][i], out[False][i]) < tol, (i, rel(out[True][i], out[False][i]))
