"""The ``torch.library`` boundary (north_star: "a thin C-ABI torch.library extension"; library.py): every kernel
launcher of include/tss_b200.h is a registered ``tss_b200`` operator whose schema mirrors the C prototype, with a fake
implementation; the functional operators carry autograd formulas.  CPU part: registration, schemas, FakeTensorMode
(no kernel runs).  GPU part: ``torch.library.opcheck`` and gradients against stock torch ops."""
import pytest
import torch

from torch_semantic_segmentation_b200 import _lib, library, ops


def test_every_launcher_of_the_header_is_a_registered_operator():
    protos = _lib.parse_header(with_const=True)
    launchers = [n for n, (ret, ps) in protos.items() if ret == 'int' and any(p[0] == 'stream' for p in ps)]
    assert len(launchers) >= 55 and set(launchers) == set(library.OPS)
    for name in launchers:
        op, names = library.OPS[name]
        schema = op._schema
        assert schema.name == 'tss_b200::' + name[4:]
        assert [a.name for a in schema.arguments] == names == [p[0] for p in protos[name][1] if p[0] != 'stream']
        for arg, (pname, kind, base, const) in zip(schema.arguments, [p for p in protos[name][1] if p[0] != 'stream']):
            if kind == 'ptr':
                assert str(arg.type) == 'Optional[Tensor]'
                mutated = arg.alias_info is not None and arg.alias_info.is_write
                assert mutated == (not const and pname not in library.HOST_ARRAYS), (name, pname)
            else:
                assert str(arg.type) == ('float' if base in ('float', 'double') else 'int')
        assert len(schema.returns) == 0


def test_calls_reach_the_backend_through_the_dispatcher(fake_backend):
    seen = []
    real = fake_backend.call
    fake_backend.call = lambda name, kw: (seen.append(name), real(name, kw))[1]
    try:
        x = ops.empty_nhwc(1, 8, 4, 4, torch.float32, 'cpu').normal_()
        ops.dwconv_fwd(x, torch.randn(8, 1, 3, 3), 1, 1)
        torch.ops.tss_b200.bilinear(x, 8, 8)
    finally:
        fake_backend.call = real
    assert seen == ['tss_dwconv3x3_fwd', 'tss_bilinear_fwd']


def test_fake_tensor_mode_runs_no_kernel_and_gets_the_layouts_right(fake_backend):
    from torch._subclasses.fake_tensor import FakeTensorMode
    calls = []
    real = fake_backend.call
    fake_backend.call = lambda name, kw: (calls.append(name), real(name, kw))[1]
    try:
        with FakeTensorMode():
            x = ops.empty_nhwc(2, 16, 12, 20, torch.float32, 'cpu')
            y = torch.ops.tss_b200.dwconv3x3(x, torch.empty(16, 1, 3, 3), 2, 1)
            assert tuple(y.shape) == (2, 16, 6, 10) and ops.geom(y) == (2, 16, 6, 10, 16)
            s = torch.ops.tss_b200.pwconv(y, torch.empty(19, 16, 1, 1), torch.empty(19))
            assert tuple(s.shape) == (2, 19, 6, 10) and ops.geom(s)[4] == 32            # 19 classes in a 32-channel pitch
            up = torch.ops.tss_b200.upsample_logits(s, 48, 80)
            assert tuple(up.shape) == (2, 19, 48, 80) and up.is_contiguous()
            loss, dlogits = torch.ops.tss_b200.cross_entropy(up, torch.empty(2, 48, 80, dtype=torch.int64), 255)
            assert loss.shape == () and dlogits.shape == up.shape
            # an out-variant launcher under fake tensors: nothing happens, nothing is launched
            torch.ops.tss_b200.bn_fold(None, None, torch.empty(8), torch.empty(8), 1e-5, torch.empty(8), torch.empty(8), 8)
    finally:
        fake_backend.call = real
    assert calls == []


def test_functional_operators_differentiate_like_stock_torch(fake_backend):
    g = torch.Generator().manual_seed(3)
    x = ops.empty_nhwc(2, 16, 9, 13, torch.float32, 'cpu').copy_(torch.randn(2, 16, 9, 13, generator=g)).requires_grad_(True)
    wd = torch.randn(16, 1, 3, 3, generator=g, requires_grad=True)
    wp = torch.randn(19, 16, 1, 1, generator=g, requires_grad=True)
    b = torch.randn(19, generator=g, requires_grad=True)
    t = torch.randint(0, 19, (2, 18, 26), generator=g)
    t[0, :3] = 255

    def ours():
        y = torch.ops.tss_b200.dwconv3x3(x, wd, 1, 1)
        s = torch.ops.tss_b200.pwconv(y, wp, b)
        up = torch.ops.tss_b200.bilinear(s, 18, 26)
        return torch.ops.tss_b200.cross_entropy(up.contiguous(), t, 255)[0]

    def stock():
        y = torch.nn.functional.conv2d(x, wd, None, 1, 1, 1, 16)
        s = torch.nn.functional.conv2d(y, wp, b)
        up = torch.nn.functional.interpolate(s, size=(18, 26), mode='bilinear', align_corners=True)
        return torch.nn.functional.cross_entropy(up, t, ignore_index=255)
    ga = torch.autograd.grad(ours(), [x, wd, wp, b])
    gb = torch.autograd.grad(stock(), [x, wd, wp, b])
    for a, r in zip(ga, gb):
        assert float((a - r).norm() / r.norm()) < 1e-5


@pytest.mark.gpu
def test_opcheck_on_the_gpu():
    g = torch.Generator().manual_seed(4)
    x = ops.empty_nhwc(2, 16, 10, 12, torch.float32, 'cuda').copy_(torch.randn(2, 16, 10, 12, generator=g).cuda())
    wd = torch.randn(16, 1, 3, 3, generator=g).cuda()
    wp = torch.randn(19, 16, 1, 1, generator=g).cuda()
    b = torch.randn(19, generator=g).cuda()
    t = torch.randint(0, 19, (2, 10, 12), generator=g).cuda()
    tests = ('test_schema', 'test_faketensor', 'test_autograd_registration')
    torch.library.opcheck(torch.ops.tss_b200.dwconv3x3.default, (x.requires_grad_(True), wd.requires_grad_(True), 2, 1), test_utils=tests)
    torch.library.opcheck(torch.ops.tss_b200.pwconv.default, (x, wp.requires_grad_(True), b.requires_grad_(True)), test_utils=tests)
    torch.library.opcheck(torch.ops.tss_b200.bilinear.default, (x, 20, 24), test_utils=tests)
    s = torch.ops.tss_b200.pwconv(x, wp, b).detach().requires_grad_(True)
    torch.library.opcheck(torch.ops.tss_b200.upsample_logits.default, (s, 20, 24), test_utils=tests)
    logits = torch.randn(2, 19, 10, 12, generator=g).cuda().requires_grad_(True)
    torch.library.opcheck(torch.ops.tss_b200.cross_entropy.default, (logits, t, 255), test_utils=tests)
    # out-variant launcher: schema (mutation annotations) against what the kernel really writes
    y = ops.empty_nhwc(2, 16, 5, 6, torch.float32, 'cuda')
    torch.library.opcheck(torch.ops.tss_b200.dwconv3x3_fwd.default,
                          (x.detach(), wd.detach(), y, 2, 10, 12, 16, 2, 1, None, None, 0, None, 0), test_utils=('test_schema', 'test_faketensor'))


@pytest.mark.gpu
def test_functional_operators_match_stock_torch_on_the_gpu():
    torch.backends.cudnn.allow_tf32 = False          # the yardstick must be true fp32 (cuDNN convolutions default to TF32)
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    x = ops.empty_nhwc(2, 32, 24, 40, torch.float32, 'cuda').copy_(torch.randn(2, 32, 24, 40, generator=g).cuda()).requires_grad_(True)
    wd = torch.randn(32, 1, 3, 3, generator=g).cuda().requires_grad_(True)
    wp = (torch.randn(19, 32, 1, 1, generator=g) / 6).cuda().requires_grad_(True)
    b = torch.randn(19, generator=g).cuda().requires_grad_(True)
    t = torch.randint(0, 19, (2, 96, 160), generator=g)
    t[0, :7] = 255
    t = t.cuda()
    s = torch.ops.tss_b200.pwconv(torch.ops.tss_b200.dwconv3x3(x, wd, 2, 1), wp, b)
    loss = torch.ops.tss_b200.cross_entropy(torch.ops.tss_b200.upsample_logits(s, 96, 160), t, 255)[0]
    ga = torch.autograd.grad(loss, [x, wd, wp, b])
    F = torch.nn.functional
    ref = F.cross_entropy(F.interpolate(F.conv2d(F.conv2d(x, wd, None, 2, 1, 1, 32), wp, b), size=(96, 160), mode='bilinear',
                                        align_corners=True), t, ignore_index=255)
    gb = torch.autograd.grad(ref, [x, wd, wp, b])
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    for a, r in zip(ga, gb):
        assert float((a - r).norm() / r.norm()) < 1e-4
