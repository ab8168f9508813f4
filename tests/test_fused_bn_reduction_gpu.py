"""The BatchNorm-backward reduction fused into the consumer's dgrad epilogue (csrc/pwconv_tc_bnred.cu,
csrc/dwconv_bnred.cu): kernel parity against stock torch ops, and the whole training step with and
without the fusion."""
import math

import pytest
import torch

from tests.fake_backend import FakeBackend
from torch_semantic_segmentation_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(N, C, H, W, g, dtype=torch.bfloat16):
    base = torch.randn(N, H, W, C, generator=g).to(dtype)
    return base.permute(0, 3, 1, 2), base.cuda().permute(0, 3, 1, 2)


@pytest.mark.parametrize('K,Nc,M,relu', [(384, 64, 2304, 1), (64, 128, 1000, 1), (576, 96, 577, 0), (128, 128, 9216, 1), (48, 64, 300, 1)])
def test_pw_dgrad_with_fused_bn_reduction(K, Nc, M, relu):
    g = torch.Generator().manual_seed(K + Nc + M)
    dyc, dyg = nhwc(1, Nc, 1, M, g)
    ypc, ypg = nhwc(1, K, 1, M, g)
    w = (torch.randn(Nc, K, 1, 1, generator=g) / math.sqrt(Nc)).to(torch.bfloat16).float()
    _, wpT = ops.pack_weights_bf16(w.cuda())
    mean, rstd = torch.randn(K, generator=g) * 0.2, torch.rand(K, generator=g) + 0.5
    gamma, beta = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.3
    gc, gg = nhwc(1, K, 1, M, g)
    sc, sg = torch.zeros(2 * K), torch.zeros(2 * K).cuda()
    kw = dict(M=M, K=K, Nc=Nc, lddy=Nc, ldg=K, ldyp=K, flags=relu)
    FakeBackend().call('tss_pwconv_dgrad_bnred', dict(dy=dyc, wpT=w.view(Nc, K).t().contiguous(), g=gc, yp=ypc, mean=mean, rstd=rstd,
                                                     gamma=gamma, beta=beta, sums=sc, **kw))
    _lib.backend().call('tss_pwconv_dgrad_bnred', dict(dy=dyg, wpT=wpT, g=gg, yp=ypg, mean=mean.cuda(), rstd=rstd.cuda(),
                                                      gamma=gamma.cuda(), beta=beta.cuda(), sums=sg, **kw))
    torch.cuda.synchronize()
    assert rel(gg, gc) < 5e-3, rel(gg, gc)
    assert rel(sg, sc) < 2e-3, rel(sg, sc)


@pytest.mark.parametrize('C,N,H,W,relu', [(384, 2, 12, 16, 1), (128, 1, 24, 40, 1), (576, 2, 6, 8, 0), (64, 1, 9, 13, 1), (96, 1, 5, 7, 1)])
def test_dw_dgrad_with_fused_bn_reduction(C, N, H, W, relu):
    g = torch.Generator().manual_seed(C + H)
    dyc, dyg = nhwc(N, C, H, W, g)
    ypc, ypg = nhwc(N, C, H, W, g)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    gc, gg = nhwc(N, C, H, W, g)
    sc, sg = torch.zeros(2 * C), torch.zeros(2 * C).cuda()
    kw = dict(N=N, H=H, W=W, C=C, flags=relu, dtype=1)
    FakeBackend().call('tss_dwconv3x3_dgrad_bnred', dict(dy=dyc, w=w, g=gc, yp=ypc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sc, **kw))
    _lib.backend().call('tss_dwconv3x3_dgrad_bnred', dict(dy=dyg, w=w.cuda(), g=gg, yp=ypg, mean=mean.cuda(), rstd=rstd.cuda(),
                                                         gamma=gamma.cuda(), beta=beta.cuda(), sums=sg, **kw))
    torch.cuda.synchronize()
    assert rel(gg, gc) < 5e-3, rel(gg, gc)
    assert rel(sg, sc) < 2e-3, rel(sg, sc)


def test_training_step_with_fused_bn_reduction_matches_unfused():
    from oracle.golden_inputs import train_batch
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    x, y = train_batch('fastscnn')
    grads = {}
    keep = Fn.FUSE_BNRED
    for flag in (False, True):
        Fn.FUSE_BNRED = flag
        try:
            torch.manual_seed(0)
            model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16).train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            before = _lib.launch_count()
            CrossEntropyLoss(ignore_index=255)(model(x.cuda()), y.cuda()).backward()
            torch.cuda.synchronize()
            grads[flag] = ({k: p.grad.clone() for k, p in model.named_parameters()}, _lib.launch_count() - before)
        finally:
            Fn.FUSE_BNRED = keep
    assert grads[True][1] < grads[False][1]                       # fewer launches: the fused reductions ran
    for k in ('classifier.3.weight', 'features.0.1.conv1.0.weight', 'features.2.2.conv2.0.weight', 'downsample.0.0.weight'):
        assert rel(grads[True][0][k], grads[False][0][k]) < 3e-2, k
