"""The fused depthwise + pointwise inference kernel (csrc/dwpw_tc.cu): kernel parity against stock torch
ops on bf16-rounded operands, and whole-model parity of the fused against the two-kernel eval path."""
import math

import pytest
import torch

from tests.fake_backend import FakeBackend
from torch_semantic_segmentation_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize('N,C,H,W,Nc,use_res,relu1,relu2', [
    (2, 384, 12, 16, 64, True, True, True), (1, 128, 24, 40, 128, False, False, True), (2, 576, 6, 8, 96, True, True, True),
    (1, 64, 9, 13, 16, False, True, False), (1, 768, 3, 5, 128, True, True, True), (3, 128, 17, 33, 48, False, False, False)])
def test_fused_dw_pw_inference(N, C, H, W, Nc, use_res, relu1, relu2):
    g = torch.Generator().manual_seed(N * C + H)
    dt = torch.bfloat16
    xb = torch.randn(N, H, W, C, generator=g).to(dt)
    xc, xg = xb.permute(0, 3, 1, 2), xb.cuda().permute(0, 3, 1, 2)
    w_dw = torch.randn(C, 1, 3, 3, generator=g) / 3
    w_pw = (torch.randn(Nc, C, 1, 1, generator=g) / math.sqrt(C)).to(dt).float()
    s1, b1 = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    s2, b2 = torch.rand(Nc, generator=g) + 0.5, torch.randn(Nc, generator=g) * 0.1
    rb = torch.randn(N, H, W, Nc, generator=g).to(dt)
    rc, rg = (rb.permute(0, 3, 1, 2), rb.cuda().permute(0, 3, 1, 2)) if use_res else (None, None)
    wp, _ = ops.pack_weights_bf16(w_pw.cuda())
    yc = torch.empty(N, H, W, Nc, dtype=dt).permute(0, 3, 1, 2)
    yg = torch.empty(N, H, W, Nc, dtype=dt, device='cuda').permute(0, 3, 1, 2)
    kw = dict(flags1=int(relu1), N=N, H=H, W=W, C=C, Nc=Nc, ldy=Nc, ldr=Nc if use_res else 0, flags2=int(relu2))
    FakeBackend().call('tss_dwpw_fwd', dict(x=xc, w_dw=w_dw, scale1=s1, shift1=b1, wp=w_pw.view(Nc, C), y=yc, scale2=s2,
                                            shift2=b2, res=rc, **kw))
    _lib.backend().call('tss_dwpw_fwd', dict(x=xg, w_dw=w_dw.cuda(), scale1=s1.cuda(), shift1=b1.cuda(), wp=wp, y=yg,
                                             scale2=s2.cuda(), shift2=b2.cuda(), res=rg, **kw))
    torch.cuda.synchronize()
    assert rel(yg, yc) < 1e-2, rel(yg, yc)


def test_fastscnn_inference_with_fused_blocks():
    from oracle.golden_inputs import eval_input
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.nn import blocks
    torch.manual_seed(0)
    model = fastscnn(3, 19).cuda().set_compute_dtype(torch.bfloat16).eval()
    x = eval_input('fastscnn').cuda()
    with torch.no_grad():
        keep = (blocks.FUSE_DW_PW, blocks.FUSE_MIN_TILES)
        try:
            blocks.FUSE_DW_PW = False
            ref = model(x)
            before = _lib.launch_count()
            model(x)
            plain = _lib.launch_count() - before
            blocks.FUSE_DW_PW, blocks.FUSE_MIN_TILES = True, 0       # force the fused path on this small input
            out = model(x)
            before = _lib.launch_count()
            model(x)
            fused = _lib.launch_count() - before
        finally:
            blocks.FUSE_DW_PW, blocks.FUSE_MIN_TILES = keep
    torch.cuda.synchronize()
    assert fused < plain
    assert rel(out, ref) < 2e-2, rel(out, ref)
