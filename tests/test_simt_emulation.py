"""The plain SIMT kernels that have not run on a B200 yet (csrc/ppm.cu, csrc/augment.cu, the stride-2 kernel of
csrc/dwconv_bnred.cu), executed on the CPU by
tests/simt_emu/: the .cu files are compiled for the host with g++ against a small emulation of the CUDA subset
they use (one OS thread per CUDA thread, barriers for __syncthreads, slot exchange for warp shuffles), and driven
through the SAME C ABI and the same Python host code as on the GPU.  This checks the kernels' index arithmetic,
shared-memory choreography, launch configurations and the address-table plumbing against the torch emulation
of tests/fake_backend.py and the oracle -- everything except the hardware itself."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from tests.fake_backend import FakeBackend
from torch_semantic_segmentation_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'torch_semantic_segmentation_b200', 'csrc')
EMU = os.path.join(ROOT, 'tests', 'simt_emu')
FULL_EXCLUDE = ('pwconv_tc.cu', 'dwpw_tc.cu', 'api.cu')       # raw-asm tcgen05 GEMMs not converted yet; api.cu's role is emu_runtime.cpp
SOURCES = ['ppm.cu', 'augment.cu', 'dwconv_bnred.cu',        # dwconv_bnred.cu: the stride-2 (plain SIMT) kernel only
           'bn_fused.cu', 'pwconv_tc_bwd.cu', 'stem_tc.cu', 'dwconv_bwd_fused.cu', 'metrics.cu', 'dropout.cu', 'dwconv_bnin.cu', 'pwconv_tc_fwd_bnin.cu',
           'pwconv_tc_bnred.cu', 'dwconv.cu', 'dwconv_tma.cu', 'bn.cu']       # validated on the B200: calibrate the emulation itself                               # on the functional tcgen05/TMA/mbarrier emulation


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


class EmulatedBackend:
    """Entry points exported by the host build run the real kernel code; everything else is the torch emulation."""

    def __init__(self, so_path):
        self.lib = ctypes.CDLL(so_path)
        self.protos = _lib.parse_header()
        self.fake = FakeBackend()
        self.emulated_calls = 0
        self.emulated_names = {}

    def call(self, name, kwargs):
        fn = getattr(self.lib, name, None)
        if fn is None:
            return self.fake.call(name, kwargs)
        ret, params = self.protos[name]
        args, keep = [], []
        for pname, kind, base in params:
            if pname == 'stream':
                args.append(ctypes.c_void_p(0))
                continue
            v = kwargs[pname]
            if kind == 'ptr':
                if v is None:
                    args.append(ctypes.c_void_p(0))
                elif hasattr(v, 'array'):
                    args.append(ctypes.cast(v.array, ctypes.c_void_p))
                else:
                    assert isinstance(v, torch.Tensor) and v.device.type == 'cpu'
                    want = _lib._PTR_DTYPE[base]
                    assert want is None or v.dtype == want, (name, pname, v.dtype)
                    keep.append(v)
                    args.append(ctypes.c_void_p(v.data_ptr()))
            else:
                args.append(_lib._CTYPES[base](v))
        fn.restype = ctypes.c_int
        rc = fn(*args)
        self.emulated_calls += 1
        self.emulated_names[name] = self.emulated_names.get(name, 0) + 1
        if rc != 0:
            self.lib.tss_last_error.restype = ctypes.c_char_p
            raise RuntimeError('%s failed (%d): %s' % (name, rc, self.lib.tss_last_error().decode()))
        return rc


@pytest.fixture(scope='module')
def emulated(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('simt_emu') / 'host_emu.so')
    cmd = ['g++', '-std=c++20', '-O1', '-ffp-contract=off', '-DTSS_HOST_EMU', '-x', 'c++', '-shared', '-fPIC', '-pthread',
           '-I', EMU, '-I', CSRC] + [os.path.join(CSRC, f) for f in SOURCES] + [os.path.join(EMU, 'emu_runtime.cpp'), '-o', so]
    subprocess.check_call(cmd)
    prev = _lib._backend
    be = EmulatedBackend(so)
    yield be
    _lib.set_backend(prev)


def _ppm_run(backend, flag, N, H, W, dtype):
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models.fastscnn import PyramidPoolingModule
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    _lib.set_backend(backend)
    keep = Fn.FUSE_PPM
    Fn.FUSE_PPM = flag
    try:
        torch.manual_seed(0)
        m = set_compute_dtype(PyramidPoolingModule(128, 128), dtype, pw_impl=0).train()
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.add_(0.3 * torch.randn(p.shape))          # non-trivial gamma / beta
        FlatAdamW(m.parameters(), lr=1e-3).zero_grad()
        g = torch.Generator().manual_seed(1)
        x = ops.as_nhwc(torch.randn(N, 128, H, W, generator=g).to(dtype)).requires_grad_()
        out = m(x)
        (out.float() * torch.randn(out.shape, generator=g)).sum().backward()
        return (out.detach().float(), x.grad.float(), {k: p.grad.clone() for k, p in m.named_parameters()},
                {k: v.clone().float() for k, v in m.state_dict().items() if 'running' in k or 'tracked' in k})
    finally:
        Fn.FUSE_PPM = keep


@pytest.mark.parametrize('shape,dtype', [((3, 6, 9), torch.float32), ((2, 8, 8), torch.bfloat16), ((20, 1, 2), torch.float32)])
def test_grouped_pyramid_kernels_on_the_simt_emulation(emulated, shape, dtype):
    N, H, W = shape
    grouped = lambda: sum(v for k, v in emulated.emulated_names.items() if k.startswith('tss_ppm'))
    before = grouped()
    got = _ppm_run(emulated, True, N, H, W, dtype)
    assert grouped() - before == 4          # branches fwd, concat fwd, concat bwd, branches bwd
    want = _ppm_run(FakeBackend(), False, N, H, W, dtype)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert rel(got[0], want[0]) < tol, ('out', rel(got[0], want[0]))
    assert rel(got[1], want[1]) < 3 * tol, ('dx', rel(got[1], want[1]))
    for k in want[2]:
        assert rel(got[2][k], want[2][k]) < 5 * tol, (k, rel(got[2][k], want[2][k]))
    for k in want[3]:
        assert rel(got[3][k], want[3][k]) < tol, (k, rel(got[3][k], want[3][k]))


def test_input_pipeline_kernel_on_the_simt_emulation(emulated):
    from oracle import augment as A
    from torch_semantic_segmentation_b200.data import DeviceTransform, eval_transform
    _lib.set_backend(emulated)
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'augment.npz'))
    before = emulated.emulated_calls
    for seed, h, w, scale, hf, wf, flip, crop in A.GOLDEN_CASES:
        img, lab = A.sample(seed, h, w)
        x, y = DeviceTransform(crop=crop)(torch.from_numpy(img)[None], torch.from_numpy(lab)[None],
                                           draws=[(scale, hf, wf, bool(flip))])
        np.testing.assert_array_equal(x[0].numpy(), gold['image_%d' % seed])
        np.testing.assert_array_equal(y[0].numpy(), gold['label_%d' % seed].astype(np.int64))
    samples = [A.sample(30 + i, 48, 80) for i in range(3)]          # a batch with a different draw per sample
    images = torch.from_numpy(np.stack([s[0] for s in samples]))
    labels = torch.from_numpy(np.stack([s[1] for s in samples]))
    draws = [(1.5, 0.2, 0.7, True), (2.2, 0.9, 0.1, False), (3.0, 0.5, 0.5, True)]
    x, y = DeviceTransform(crop=(40, 64))(images, labels, draws=draws)
    for i, d in enumerate(draws):
        ex, ey = A.train_transform(samples[i][0], samples[i][1], d[0], d[1], d[2], d[3], (40, 64))
        np.testing.assert_array_equal(x[i].numpy(), ex)
        np.testing.assert_array_equal(y[i].numpy(), ey)
    x, y = eval_transform()(images, labels)
    ex, ey = A.eval_transform(*samples[2])
    np.testing.assert_array_equal(x[2].numpy(), ex)
    np.testing.assert_array_equal(y[2].numpy(), ey)
    assert emulated.emulated_calls - before == len(A.GOLDEN_CASES) + 2


def _nhwc(N, C, H, W, g, dtype):
    return torch.randn(N, H, W, C, generator=g).to(dtype).permute(0, 3, 1, 2)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,Hi,Wi,relu', [(32, 2, 12, 16, 1), (48, 1, 9, 13, 0), (64, 3, 7, 5, 1), (8, 1, 1, 1, 1), (384, 1, 4, 6, 1)])
def test_stride2_dgrad_with_fused_reduction_on_the_simt_emulation(emulated, C, N, Hi, Wi, relu, dtype):
    g = torch.Generator().manual_seed(C + Hi)
    Ho, Wo = (Hi - 1) // 2 + 1, (Wi - 1) // 2 + 1
    dy, yp = _nhwc(N, C, Ho, Wo, g, dtype), _nhwc(N, C, Hi, Wi, g, dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        gout = torch.zeros(N, Hi, Wi, C, dtype=dtype).permute(0, 3, 1, 2)
        sums = torch.zeros(2 * C)
        be.call('tss_dwconv3x3_dgrad_s2_bnred', dict(dy=dy, w=w, g=gout, N=N, Hi=Hi, Wi=Wi, C=C, yp=yp, mean=mean, rstd=rstd,
                                                     gamma=gamma, beta=beta, flags=relu, sums=sums, dtype=_lib.dtype_code(dtype)))
        outs[name] = (gout.float(), sums)
    tol = 1e-5 if dtype == torch.float32 else 5e-3
    assert rel(outs['emu'][0], outs['ref'][0]) < tol
    assert rel(outs['emu'][1], outs['ref'][1]) < 2e-3


@pytest.mark.parametrize('link', [False, True])
@pytest.mark.parametrize('M_shape,K,Nc,relu', [((2, 9, 13), 64, 384, 1), ((1, 16, 8), 128, 128, 1), ((2, 5, 7), 96, 576, 0),
                                               ((3, 8, 8), 32, 48, 1), ((1, 1, 3), 16, 8, 1), ((1, 12, 25), 48, 64, 1),
                                               ((1, 4, 4), 128, 768, 1)])
def test_pw_backward_with_bn_apply_on_the_tcgen05_emulation(emulated, M_shape, K, Nc, relu, link):
    """csrc/pwconv_tc_bwd.cu through tests/simt_emu/tcgen05_emu.h: warp roles, mbarrier protocol (a deadlock shows
    up as a test timeout), the thread-built swizzled A tile, channel / row tails, the fused producer reduction."""
    N, H, W = M_shape
    M = N * H * W
    g = torch.Generator().manual_seed(K + Nc + M)
    dt = torch.bfloat16
    dz, y, yp = _nhwc(N, Nc, H, W, g, dt), _nhwc(N, Nc, H, W, g, dt), _nhwc(N, K, H, W, g, dt)
    par = lambda C: (torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) + 0.5,
                     torch.randn(C, generator=g) * 0.3)
    mean, rstd, gamma, beta = par(Nc)
    pmean, prstd, pgamma, pbeta = par(K)
    wpT = (torch.randn(K, Nc, generator=g) / Nc ** 0.5).to(dt)
    fake = FakeBackend()
    sums = torch.zeros(2 * Nc)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=Nc,
                                        lddz=Nc, ldz=0, ldy=Nc, flags=relu, dtype=1))
    outs = {}
    for name, be in (('ref', fake), ('emu', emulated)):
        dy = torch.zeros(N, H, W, Nc, dtype=dt).permute(0, 3, 1, 2)
        dx = torch.zeros(N, H, W, K, dtype=dt).permute(0, 3, 1, 2)
        dgamma, dbeta, psums = torch.ones(Nc), torch.ones(Nc), torch.zeros(2 * K)
        be.call('tss_pwconv_bwd_fused', dict(
            dz=dz, y=y, lddz=Nc, ldy=Nc, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu, count=M, dy=dy,
            lddy=Nc, dgamma=dgamma, dbeta=dbeta, wpT=wpT, dx=dx, M=M, K=K, Nc=Nc, lddx=K, yp=yp if link else None,
            ldyp=K if link else 0, pmean=pmean if link else None, prstd=prstd if link else None,
            pgamma=pgamma if link else None, pbeta=pbeta if link else None, pflags=1 if link else 0,
            psums=psums if link else None))
        outs[name] = (dy.float(), dx.float(), dgamma, dbeta, psums)
    r, e = outs['ref'], outs['emu']
    assert rel(e[0], r[0]) < 5e-3, ('dy', rel(e[0], r[0]))
    assert rel(e[1], r[1]) < 8e-3, ('dx', rel(e[1], r[1]))
    assert rel(e[2], r[2]) < 1e-6 and rel(e[3], r[3]) < 1e-6
    if link:
        assert rel(e[4], r[4]) < 5e-3, ('psums', rel(e[4], r[4]))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('N,C,H,W,relu,with_res', [(2, 32, 9, 7, 1, 0), (1, 384, 5, 3, 1, 1), (3, 48, 16, 24, 0, 1), (2, 8, 1, 1, 1, 0)])
def test_finalize_folded_into_apply_on_the_simt_emulation(emulated, N, C, H, W, relu, with_res, dtype):
    g = torch.Generator().manual_seed(C + H)
    y = _nhwc(N, C, H, W, g, dtype)
    res = _nhwc(N, C, H, W, g, dtype) if with_res else None
    M = N * H * W
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        scratch = torch.zeros(3 * C, dtype=torch.float64)
        scratch[:C] = y.double().sum(dim=(0, 2, 3))
        scratch[C:2 * C] = (y.double() ** 2).sum(dim=(0, 2, 3))
        scratch[2 * C:] = 7.0                                        # stale backward sums: must be cleared too
        rm, rv, nbt = torch.zeros(C), torch.ones(C), torch.zeros((), dtype=torch.int64)
        mean, rstd, ticket = torch.zeros(C), torch.zeros(C), torch.zeros(1, dtype=torch.int32)
        z = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        be.call('tss_bn_finalize_apply', dict(stats=scratch, count=M, gamma=gamma, beta=beta, running_mean=rm, running_var=rv,
                                              num_batches_tracked=nbt, momentum=0.1, eps=1e-5, mean=mean, rstd=rstd, ticket=ticket,
                                              clear_n=3 * C, y=y, res=res, z=z, M=M, C=C, ldy=C, ldr=C if with_res else 0, ldz=C,
                                              flags=relu, dtype=_lib.dtype_code(dtype)))
        outs[name] = (z.float(), mean, rstd, rm, rv, int(nbt), float(scratch.abs().max()), int(ticket))
    r, e = outs['ref'], outs['emu']
    tol = 1e-6 if dtype == torch.float32 else 4e-3
    assert rel(e[0], r[0]) < tol
    for i in (1, 2, 3, 4):
        assert rel(e[i], r[i]) < 1e-6
    assert e[5] == 1 and e[6] == 0.0 and e[7] == 0              # counted once, scratch cleared, ticket reset


@pytest.mark.parametrize('N,H,W,mode', [(2, 16, 24, 'train'), (1, 32, 300, 'train'), (1, 7, 9, 'eval'), (3, 2, 2, 'train'),
                                          (1, 64, 520, 'eval')])
def test_stem_on_the_tcgen05_emulation(emulated, N, H, W, mode):
    """csrc/stem_tc.cu: thread-built implicit-GEMM operands (image patches and weights), K = 27 padded to 32,
    tiles of 128 output pixels with a row tail, statistics or the folded eval epilogue."""
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(N, 3, H, W, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) / 5
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    scale = shift = None
    if mode == 'eval':
        scale, shift = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.2
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        y = torch.zeros(N, Ho, Wo, 32, dtype=torch.bfloat16).permute(0, 3, 1, 2)
        stats = torch.zeros(64, dtype=torch.float64) if mode == 'train' else None
        be.call('tss_stem3x3s2_fwd_tc', dict(x=x, w=w, y=y, N=N, H=H, W=W, Cout=32, scale=scale, shift=shift,
                                             flags=1 if mode == 'eval' else 0, stats=stats))
        outs[name] = (y.float(), stats)
    assert rel(outs['emu'][0], outs['ref'][0]) < 4e-3
    if mode == 'train':
        assert rel(outs['emu'][1], outs['ref'][1]) < 1e-5


@pytest.mark.parametrize('N,H,W', [(2, 16, 24), (1, 32, 300), (1, 7, 9), (3, 2, 2), (1, 20, 260)])
def test_stem_weight_gradient_on_the_tcgen05_emulation(emulated, N, H, W):
    """csrc/stem_tc.cu wgrad: pixel index as the reduction dimension, thread-built transposed operands, 2-stage
    mbarrier ring over a persistent CTA's share of the k-blocks, accumulation into dw."""
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(N, 3, H, W, generator=g)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dy = _nhwc(N, 32, Ho, Wo, g, torch.bfloat16)
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        dw = torch.full((32, 3, 3, 3), 0.5)                          # accumulates onto what is there
        be.call('tss_stem3x3s2_wgrad_tc', dict(x=x, dy=dy, dw=dw, N=N, H=H, W=W, Cout=32))
        outs[name] = dw
    assert rel(outs['emu'] - 0.5, outs['ref'] - 0.5) < 2e-3


@pytest.mark.parametrize('M_shape,K,Nc,relu', [((2, 9, 13), 64, 384, 1), ((1, 16, 8), 128, 128, 1), ((2, 5, 7), 96, 576, 0),
                                               ((3, 8, 8), 32, 48, 1), ((1, 12, 25), 48, 64, 1)])
def test_emulation_agrees_with_a_kernel_validated_on_the_gpu(emulated, M_shape, K, Nc, relu):
    """csrc/pwconv_tc_bnred.cu is part of the default path and parity-green on the B200
    (tests/test_fused_bn_reduction_gpu.py).  Its host build must give the same answers on the emulation: this pins
    the emulation's model of TMA swizzling, UMMA descriptors, TMEM addressing and mbarrier phases to a kernel
    whose behaviour on the hardware is known."""
    N, H, W = M_shape
    M = N * H * W
    g = torch.Generator().manual_seed(K + Nc + M)
    dt = torch.bfloat16
    dy, yp = _nhwc(N, Nc, H, W, g, dt), _nhwc(N, K, H, W, g, dt)
    mean, rstd = torch.randn(K, generator=g) * 0.2, torch.rand(K, generator=g) + 0.5
    gamma, beta = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.3
    wpT = (torch.randn(K, Nc, generator=g) / Nc ** 0.5).to(dt)
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        gout = torch.zeros(N, H, W, K, dtype=dt).permute(0, 3, 1, 2)
        sums = torch.zeros(2 * K)
        be.call('tss_pwconv_dgrad_bnred', dict(dy=dy, wpT=wpT, g=gout, M=M, K=K, Nc=Nc, lddy=Nc, ldg=K, yp=yp, ldyp=K, mean=mean,
                                               rstd=rstd, gamma=gamma, beta=beta, flags=relu, sums=sums))
        outs[name] = (gout.float(), sums)
    assert rel(outs['emu'][0], outs['ref'][0]) < 5e-3
    assert rel(outs['emu'][1], outs['ref'][1]) < 5e-3


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_grouped_pyramid_eval_on_the_simt_emulation(emulated, dtype):
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models.fastscnn import PyramidPoolingModule
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    keep = Fn.FUSE_PPM
    outs = {}
    try:
        for name, be, flag in (('ref', FakeBackend(), False), ('emu', emulated, True)):
            _lib.set_backend(be)
            Fn.FUSE_PPM = flag
            torch.manual_seed(0)
            m = set_compute_dtype(PyramidPoolingModule(128, 128), dtype, pw_impl=0)
            with torch.no_grad():
                for bn in [p[1][1] for p in m.pyramids]:
                    bn.running_mean.normal_(0, 0.2)
                    bn.running_var.uniform_(0.5, 1.5)
                    bn.weight.uniform_(0.5, 1.5)
                    bn.bias.normal_(0, 0.2)
            m.eval()
            g = torch.Generator().manual_seed(1)
            x = ops.as_nhwc(torch.randn(1, 128, 8, 16, generator=g).to(dtype))
            grouped = lambda: sum(v for k, v in getattr(be, 'emulated_names', {}).items() if k.startswith('tss_ppm'))
            before = grouped()
            with torch.no_grad():
                outs[name] = m(x).float()
            if flag:
                assert grouped() - before == 2          # branches + concat
    finally:
        Fn.FUSE_PPM = keep
    assert rel(outs['emu'], outs['ref']) < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W,stride,dil', [(64, 2, 12, 20, 1, 1), (48, 1, 9, 13, 1, 1), (384, 1, 5, 7, 1, 1), (32, 1, 16, 24, 2, 1),
                                                (128, 1, 12, 12, 1, 4)])
def test_emulation_agrees_with_the_validated_depthwise_kernels(emulated, C, N, H, W, stride, dil, dtype):
    """csrc/dwconv_tma.cu / dwconv.cu are parity-green on the B200 (tests/test_kernels_gpu.py): forward with
    statistics, dgrad, wgrad and the stride-1 dgrad with fused reduction must give the same answers on the
    emulation -- this pins its model of 4-D TMA boxes (halo coordinates, zero fill) and of the mbarrier pipelines."""
    g = torch.Generator().manual_seed(C + H + stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x, dy = _nhwc(N, C, H, W, g, dtype), _nhwc(N, C, Ho, Wo, g, dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    code = _lib.dtype_code(dtype)
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        y = torch.zeros(N, Ho, Wo, C, dtype=dtype).permute(0, 3, 1, 2)
        stats = torch.zeros(2 * C, dtype=torch.float64)
        be.call('tss_dwconv3x3_fwd', dict(x=x, w=w, y=y, N=N, Hi=H, Wi=W, C=C, stride=stride, dilation=dil, scale=None, shift=None,
                                          flags=0, stats=stats, dtype=code))
        dx = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        be.call('tss_dwconv3x3_dgrad', dict(dy=dy, w=w, dx=dx, N=N, Hi=H, Wi=W, C=C, stride=stride, dilation=dil, dtype=code))
        dw = torch.zeros(C, 1, 3, 3)
        be.call('tss_dwconv3x3_wgrad', dict(x=x, dy=dy, dw=dw, N=N, Hi=H, Wi=W, C=C, stride=stride, dilation=dil, dtype=code))
        outs[name] = [y.float(), stats, dx.float(), dw]
        if stride == 1 and dil == 1 and C % 32 == 0:
            mean, rstd = torch.linspace(-0.2, 0.2, C), torch.linspace(0.5, 1.5, C)
            gamma, beta = torch.linspace(0.5, 1.5, C), torch.linspace(-0.3, 0.3, C)
            gout = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
            sums = torch.zeros(2 * C)
            be.call('tss_dwconv3x3_dgrad_bnred', dict(dy=dy, w=w, g=gout, N=N, H=H, W=W, C=C, yp=x, mean=mean, rstd=rstd, gamma=gamma,
                                                      beta=beta, flags=1, sums=sums, dtype=code))
            outs[name] += [gout.float(), sums]
    tol = 1e-5 if dtype == torch.float32 else 5e-3
    for a, b in zip(outs['emu'], outs['ref']):
        assert rel(a, b) < max(tol, 2e-3 if a.dtype == torch.float32 and a.dim() == 1 else tol), rel(a, b)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W', [(128, 2, 12, 20), (96, 1, 9, 40)])
def test_depthwise_backward_with_bn_apply_without_a_producer(emulated, C, N, H, W, dtype):
    g = torch.Generator().manual_seed(C + H)
    dz, y = _nhwc(N, C, H, W, g, dtype), _nhwc(N, C, H, W, g, dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    M, code = N * H * W, _lib.dtype_code(dtype)
    fake = FakeBackend()
    sums = torch.zeros(2 * C)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=C,
                                        lddz=C, ldz=0, ldy=C, flags=1, dtype=code))
    outs = {}
    for name, be in (('ref', fake), ('emu', emulated)):
        dy = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        gout = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        be.call('tss_dwconv3x3_bwd_fused', dict(dz=dz, y=y, w=w, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=1,
                                                count=M, dy=dy, dgamma=None, dbeta=None, g=gout, N=N, H=H, W=W, C=C, yp=None,
                                                pmean=None, prstd=None, pgamma=None, pbeta=None, pflags=0, psums=None, dtype=code))
        outs[name] = (dy.float(), gout.float())
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel(outs['emu'][0], outs['ref'][0]) < tol and rel(outs['emu'][1], outs['ref'][1]) < 2 * tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W,relu', [(64, 2, 12, 20, 1), (384, 1, 5, 7, 1), (96, 1, 9, 40, 0), (32, 3, 1, 1, 1), (128, 1, 17, 33, 1)])
def test_depthwise_backward_with_bn_apply_on_the_emulation(emulated, C, N, H, W, relu, dtype):
    """csrc/dwconv_bwd_fused.cu: two TMA halo tiles (dz, y) -> dy in shared memory (zero outside the image) ->
    flipped-tap stencil -> producer mask + reduction; dy interior stored for the weight gradient."""
    g = torch.Generator().manual_seed(C + H + W)
    dz, y, yp = _nhwc(N, C, H, W, g, dtype), _nhwc(N, C, H, W, g, dtype), _nhwc(N, C, H, W, g, dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    par = lambda: (torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) + 0.5,
                   torch.randn(C, generator=g) * 0.3)
    mean, rstd, gamma, beta = par()
    pmean, prstd, pgamma, pbeta = par()
    M = N * H * W
    fake = FakeBackend()
    sums = torch.zeros(2 * C)
    code = _lib.dtype_code(dtype)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=C,
                                        lddz=C, ldz=0, ldy=C, flags=relu, dtype=code))
    outs = {}
    for name, be in (('ref', fake), ('emu', emulated)):
        dy = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        gout = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        dgamma, dbeta, psums = torch.ones(C), torch.ones(C), torch.zeros(2 * C)
        be.call('tss_dwconv3x3_bwd_fused', dict(dz=dz, y=y, w=w, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu,
                                                count=M, dy=dy, dgamma=dgamma, dbeta=dbeta, g=gout, N=N, H=H, W=W, C=C, yp=yp,
                                                pmean=pmean, prstd=prstd, pgamma=pgamma, pbeta=pbeta, pflags=1, psums=psums,
                                                dtype=code))
        outs[name] = (dy.float(), gout.float(), dgamma, dbeta, psums)
    r, e = outs['ref'], outs['emu']
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel(e[0], r[0]) < tol, ('dy', rel(e[0], r[0]))
    assert rel(e[1], r[1]) < 2 * tol, ('g', rel(e[1], r[1]))
    assert rel(e[2], r[2]) < 1e-6 and rel(e[3], r[3]) < 1e-6
    assert rel(e[4], r[4]) < max(tol, 2e-3), ('psums', rel(e[4], r[4]))


@pytest.mark.parametrize('variant', ['0', '1'])
@pytest.mark.parametrize('kind', ['random', 'piecewise', 'odd'])
def test_confusion_matrix_kernels_on_the_simt_emulation(emulated, monkeypatch, variant, kind):
    """csrc/metrics.cu: the default kernel (validated on the B200; calibrates the warp-vote emulation) and the
    warp-private variant (TSS_CM_VARIANT=1, not yet run on the GPU) are exact against the oracle's bincount."""
    from oracle import confusion as o_cm
    monkeypatch.setenv('TSS_CM_VARIANT', variant)
    g = torch.Generator().manual_seed(5)
    n = 4099 if kind == 'odd' else 8192
    if kind == 'piecewise':
        pred = torch.randint(0, 19, (n // 64,), generator=g).repeat_interleave(64)
        label = torch.randint(0, 19, (n // 128,), generator=g).repeat_interleave(128)
    else:
        pred, label = torch.randint(0, 19, (n,), generator=g), torch.randint(0, 19, (n,), generator=g)
    label = label.clone()
    label[torch.rand(n, generator=g) < 0.1] = 255
    cm = torch.zeros(19, 19, dtype=torch.int64)
    emulated.call('tss_confusion_from_labels', dict(pred=pred, target=label, n=n, C=19, cm=cm))
    want = o_cm.confusion_matrix(pred.numpy(), label.numpy(), 19)
    assert (cm.numpy() == want).all() and int(cm.sum()) == int((label != 255).sum())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_dropout_kernels_on_the_simt_emulation(emulated, dtype):
    """csrc/dropout.cu: the Philox mask is bit-identical to the numpy restatement, the backward pass regenerates the
    forward's mask from (seed, used offset), and the device-side offset advances once per forward launch."""
    from tests.philox_ref import keep_mask
    g = torch.Generator().manual_seed(3)
    N, C, H, W, p = 2, 128, 9, 7, 0.1
    x, dy = _nhwc(N, C, H, W, g, dtype), _nhwc(N, C, H, W, g, dtype)
    rng = torch.tensor([0x1234567890ABCDEF % (2 ** 62), 41, 0], dtype=torch.int64)
    n = x.numel()
    code = _lib.dtype_code(dtype)
    for step in range(2):
        y = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        dx = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        used = torch.zeros(1, dtype=torch.int64)
        emulated.call('tss_dropout_fwd', dict(x=x, y=y, n=n, p=p, rng=rng, used=used, dtype=code))
        emulated.call('tss_dropout_bwd', dict(dy=dy, dx=dx, n=n, p=p, rng=rng, used=used, dtype=code))
        assert int(used) == 41 + step and int(rng[1]) == 42 + step and int(rng[2]) == 0
        keep, scale = keep_mask(int(rng[0]), int(used), n, p)
        keep = torch.from_numpy(keep).view(N, H, W, C).permute(0, 3, 1, 2)
        want_y = torch.where(keep, x.float() * float(scale), torch.zeros(())).to(dtype)
        want_dx = torch.where(keep, dy.float() * float(scale), torch.zeros(())).to(dtype)
        assert torch.equal(y, want_y) and torch.equal(dx, want_dx)
        assert abs(float(keep.float().mean()) - (1 - p)) < 0.02


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W,stride,relu', [(64, 2, 12, 20, 1, 1), (384, 1, 5, 7, 2, 1), (96, 1, 9, 40, 1, 0), (32, 2, 16, 24, 2, 1),
                                                 (48, 1, 7, 33, 1, 1)])
def test_depthwise_kernels_with_the_input_batchnorm_folded_in(emulated, C, N, H, W, stride, relu, dtype):
    """csrc/dwconv_bnin.cu: forward (with statistics) and weight gradient on act(BN(raw input)) without materialising
    it; zero padding applies to the ACTIVATED tensor although TMA zero-fills the raw one."""
    g = torch.Generator().manual_seed(C + H + stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x, dy = _nhwc(N, C, H, W, g, dtype), _nhwc(N, C, Ho, Wo, g, dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5 + 0.3      # BN(0) != 0
    code = _lib.dtype_code(dtype)
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        y = torch.zeros(N, Ho, Wo, C, dtype=dtype).permute(0, 3, 1, 2)
        stats = torch.zeros(2 * C, dtype=torch.float64)
        be.call('tss_dwconv3x3_fwd_bnin', dict(x=x, in_scale=sc, in_shift=sh, in_flags=relu, w=w, y=y, N=N, Hi=H, Wi=W, C=C,
                                               stride=stride, stats=stats, dtype=code))
        dw = torch.zeros(C, 1, 3, 3)
        be.call('tss_dwconv3x3_wgrad_bnin', dict(x=x, in_scale=sc, in_shift=sh, in_flags=relu, dy=dy, dw=dw, N=N, Hi=H, Wi=W, C=C,
                                                 stride=stride, dtype=code))
        outs[name] = (y.float(), stats, dw)
    tol = 1e-5 if dtype == torch.float32 else 5e-3
    assert rel(outs['emu'][0], outs['ref'][0]) < tol
    assert rel(outs['emu'][1], outs['ref'][1]) < 1e-5
    assert rel(outs['emu'][2], outs['ref'][2]) < 1e-4


@pytest.mark.parametrize('M_shape,K,Nc,relu', [((2, 9, 13), 384, 64, 1), ((1, 16, 8), 128, 128, 1), ((2, 5, 7), 576, 96, 0),
                                               ((3, 8, 8), 48, 32, 1), ((1, 1, 3), 8, 16, 1), ((1, 12, 25), 768, 128, 1)])
def test_pw_forward_with_the_input_batchnorm_on_the_tcgen05_emulation(emulated, M_shape, K, Nc, relu):
    """csrc/pwconv_tc_fwd_bnin.cu: the A operand is relu(BN(raw input)) built by the threads, z stored once, the raw
    output and its statistics come from the TMEM accumulator."""
    N, H, W = M_shape
    M = N * H * W
    g = torch.Generator().manual_seed(K + Nc + M)
    dt = torch.bfloat16
    x = _nhwc(N, K, H, W, g, dt)
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.4
    wp = (torch.randn(Nc, K, generator=g) / K ** 0.5).to(dt)
    outs = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        y = torch.zeros(N, H, W, Nc, dtype=dt).permute(0, 3, 1, 2)
        z = torch.zeros(N, H, W, K, dtype=dt).permute(0, 3, 1, 2)
        stats = torch.zeros(2 * Nc, dtype=torch.float64)
        be.call('tss_pwconv_fwd_bnin', dict(x=x, ldx=K, in_scale=sc, in_shift=sh, in_flags=relu, z=z, ldz=K, wp=wp, y=y, ldy=Nc, M=M,
                                            K=K, Nc=Nc, stats=stats))
        outs[name] = (y.float(), z.float(), stats)
    assert rel(outs['emu'][1], outs['ref'][1]) < 1e-4            # z: fma vs mul+add before the bf16 rounding
    assert rel(outs['emu'][0], outs['ref'][0]) < 5e-3
    assert rel(outs['emu'][2], outs['ref'][2]) < 1e-3


@pytest.mark.parametrize('N,H,W,relu', [(2, 16, 24, 1), (1, 32, 300, 1), (1, 7, 9, 0), (3, 2, 2, 1)])
def test_stem_weight_gradient_with_bn_apply_on_the_tcgen05_emulation(emulated, N, H, W, relu):
    """csrc/stem_tc.cu, wgrad with the stem's BatchNorm-backward apply in the operand producer (no dy tensor)."""
    g = torch.Generator().manual_seed(H * W + relu)
    x = torch.randn(N, 3, H, W, generator=g)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dz, y = _nhwc(N, 32, Ho, Wo, g, torch.bfloat16), _nhwc(N, 32, Ho, Wo, g, torch.bfloat16)
    mean, rstd = torch.randn(32, generator=g) * 0.2, torch.rand(32, generator=g) + 0.5
    gamma, beta = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.3
    M = N * Ho * Wo
    fake = FakeBackend()
    sums = torch.zeros(64)
    fake.call('tss_bn_bwd_reduce', dict(dz=dz, z=None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, M=M, C=32, lddz=32,
                                        ldz=0, ldy=32, flags=relu, dtype=1))
    outs = {}
    for name, be in (('ref', fake), ('emu', emulated)):
        dw, dga, dbe = torch.full((32, 3, 3, 3), 0.25), torch.ones(32), torch.ones(32)
        be.call('tss_stem3x3s2_wgrad_tc_bn', dict(x=x, dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, flags=relu,
                                                  count=M, dw=dw, dgamma=dga, dbeta=dbe, N=N, H=H, W=W, Cout=32))
        outs[name] = (dw - 0.25, dga, dbe)
    assert rel(outs['emu'][0], outs['ref'][0]) < 5e-3
    assert rel(outs['emu'][1], outs['ref'][1]) < 1e-6 and rel(outs['emu'][2], outs['ref'][2]) < 1e-6


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W,mask', [(128, 2, 12, 20, 'y'), (96, 1, 9, 40, 'none'), (64, 3, 7, 5, 'z'), (8, 1, 1, 1, 'y'),
                                          (384, 1, 5, 7, 'y'), (48, 2, 33, 17, 'z'), (576, 1, 3, 4, 'none')])
def test_batchnorm_backward_reduction_on_the_simt_emulation(emulated, C, N, H, W, mask, dtype):
    """csrc/bn.cu, the reduction pass: (channel group, row lane) threads, U rows in flight, the three mask modes, lanes
    summed through the [PL][C] shared-memory array, pitched operands."""
    g = torch.Generator().manual_seed(C + H + W)
    pitch = C + 16
    wide = lambda: torch.randn(N, H, W, pitch, generator=g).to(dtype)
    dzb, yb, zb = wide(), wide(), wide()
    dz, y, z = (t[..., 8:8 + C].permute(0, 3, 1, 2) for t in (dzb, yb, zb))
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    M, code = N * H * W, _lib.dtype_code(dtype)
    out = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        sums = torch.full((2 * C,), 0.25)
        be.call('tss_bn_bwd_reduce', dict(dz=dz, z=z if mask == 'z' else None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta,
                                          sums=sums, M=M, C=C, lddz=pitch, ldz=pitch if mask == 'z' else 0, ldy=pitch,
                                          flags=0 if mask == 'none' else 1, dtype=code))
        out[name] = sums
    assert rel(out['emu'], out['ref']) < 2e-5, rel(out['emu'], out['ref'])


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W,mask', [(128, 2, 12, 20, 'y'), (96, 1, 9, 40, 'none'), (64, 3, 7, 5, 'z'), (8, 1, 1, 1, 'y'),
                                          (384, 1, 5, 7, 'y'), (576, 1, 3, 4, 'none'), (48, 2, 33, 17, 'z'), (96, 2, 7, 9, 'z')])
def test_batchnorm_backward_in_one_launch_on_the_simt_emulation(emulated, C, N, H, W, mask, dtype):
    """csrc/bn.cu, tss_bn_bwd_onepass: pass 1 -> (grid barrier: the emulation runs its single CTA) -> pass 2 walking the
    thread's rows backwards, all three mask modes, dres for the residual layers; against the reduce + apply pair of the
    torch emulation."""
    g = torch.Generator().manual_seed(C + H + W)
    pitch = C + 8
    wide = lambda: torch.randn(N, H, W, pitch, generator=g).to(dtype)
    dzb, yb, zb = wide(), wide(), wide()
    dz, y, z = (t[..., :C].permute(0, 3, 1, 2) for t in (dzb, yb, zb))
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    M, code = N * H * W, _lib.dtype_code(dtype)
    out = {}
    for name, be in (('ref', FakeBackend()), ('emu', emulated)):
        sums, dgamma, dbeta = torch.zeros(2 * C), torch.ones(C), torch.ones(C)
        dy = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
        dres = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2) if mask == 'z' else None
        sync = torch.zeros(4, dtype=torch.int32)
        be.call('tss_bn_bwd_onepass', dict(dz=dz, z=z if mask == 'z' else None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta,
                                           sums=sums, dy=dy, dres=dres, dgamma=dgamma, dbeta=dbeta, M=M, C=C, lddz=pitch,
                                           ldz=pitch if mask == 'z' else 0, ldy=pitch, lddy=C, lddres=C,
                                           flags=0 if mask == 'none' else 1, sync=sync, dtype=code))
        assert int(sync[2]) == 0 and int(sync[0]) == 0
        out[name] = (dy.float(), sums, dgamma, dbeta, None if dres is None else dres.float())
    r, e = out['ref'], out['emu']
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel(e[0], r[0]) < tol, ('dy', rel(e[0], r[0]))
    assert rel(e[1], r[1]) < 2e-5 and rel(e[2], r[2]) < 2e-5 and rel(e[3], r[3]) < 2e-5
    if mask == 'z':
        assert torch.equal(e[4], r[4])


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,N,H,W,mask,want_dres', [(128, 2, 12, 20, 'y', 0), (96, 1, 9, 40, 'none', 1), (64, 3, 7, 5, 'z', 1), (8, 1, 1, 1, 'y', 1),
                                                    (384, 1, 5, 7, 'y', 0), (576, 1, 3, 4, 'none', 0), (48, 2, 33, 17, 'z', 1)])
def test_batchnorm_backward_apply_on_the_simt_emulation(emulated, C, N, H, W, mask, want_dres, dtype):
    """csrc/bn.cu, the apply pass: dy = a * (g - c1 - y * k2) with the folded per-channel constants, the three mask modes,
    the optional residual-branch gradient, pitched operands; SyncBN's larger count."""
    g = torch.Generator().manual_seed(C + H + W)
    pitch = C + 8
    wide = lambda: torch.randn(N, H, W, pitch, generator=g).to(dtype)
    dzb, yb, zb = wide(), wide(), wide()
    dz, y, z = (t[..., :C].permute(0, 3, 1, 2) for t in (dzb, yb, zb))
    mean, rstd = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    sums = torch.randn(2 * C, generator=g) * 3
    M, code = N * H * W, _lib.dtype_code(dtype)
    for count in (0, 2 * M):
        out = {}
        for name, be in (('ref', FakeBackend()), ('emu', emulated)):
            dgamma, dbeta = torch.ones(C), torch.ones(C)
            dy = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2)
            dres = torch.zeros(N, H, W, C, dtype=dtype).permute(0, 3, 1, 2) if want_dres else None
            be.call('tss_bn_bwd_apply', dict(dz=dz, z=z if mask == 'z' else None, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta,
                                             sums=sums, dy=dy, dres=dres, dgamma=dgamma, dbeta=dbeta, M=M, count=count, C=C,
                                             lddz=pitch, ldz=pitch if mask == 'z' else 0, ldy=pitch, lddy=C, lddres=C,
                                             flags=0 if mask == 'none' else 1, dtype=code))
            out[name] = (dy.float(), dgamma, dbeta, None if dres is None else dres.float())
        r, e = out['ref'], out['emu']
        assert rel(e[0], r[0]) < (2e-5 if dtype == torch.float32 else 6e-3), ('dy', rel(e[0], r[0]))
        assert rel(e[1], r[1]) < 1e-6 and rel(e[2], r[2]) < 1e-6
        if want_dres:
            assert torch.equal(e[3], r[3])


@pytest.mark.skipif(os.environ.get('TSS_EMU_FULL') != '1', reason='~2 min: set TSS_EMU_FULL=1')
def test_whole_training_step_and_evaluation_on_the_emulation(tmp_path):
    """EVERY kernel call of one fp32 Fast-SCNN optimisation step (forward, fused head, backward, AdamW: ~325 calls) and
    of an evaluation forward + confusion matrix runs as real kernel code on the emulation and is compared, call by call
    on identical inputs, with the torch emulation of the ABI.  (End-to-end comparisons of this tiny net are dominated by
    the 2-sample BatchNorm of the pyramid's bin-1 branch; per-call comparisons are not.)"""
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.metrics import ConfusionMatrix
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    so = str(tmp_path / 'host_emu_all.so')
    files = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu') and f not in FULL_EXCLUDE)
    subprocess.check_call(['g++', '-std=c++20', '-O1', '-ffp-contract=off', '-DTSS_HOST_EMU', '-x', 'c++', '-shared', '-fPIC', '-pthread',
                           '-I', EMU, '-I', CSRC] + [os.path.join(CSRC, f) for f in files] + [os.path.join(EMU, 'emu_runtime.cpp'), '-o', so])
    emu, fake = EmulatedBackend(so), FakeBackend()
    inner, bad, seen = emu.call, [], set()

    def checked(name, kw):
        if 'table' in kw or not hasattr(emu.lib, name):          # address tables point at the originals / not emulated
            return inner(name, kw)
        clones = {}
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                base = v._base if v._base is not None else v
                cb = base.clone()
                clones[k] = cb.as_strided(v.shape, v.stride(), v.storage_offset()) if v._base is not None else cb
            else:
                clones[k] = v
        fake.call(name, clones)
        r = inner(name, kw)
        seen.add(name)
        for k, v in kw.items():
            if isinstance(v, torch.Tensor) and v.numel() > 0:
                if v.dim() == 1 and v.is_floating_point() and isinstance(kw.get('dz'), torch.Tensor):
                    # per-channel sums of a gradient that is mean-free in exact arithmetic are rounding noise on both sides:
                    # differences are measured against what was summed, not against the (vanishing) result
                    floor = 1e-5 * float(kw['dz'].float().abs().sum()) / max(v.numel(), 1)
                    if float((v.double() - clones[k].double()).abs().max()) < floor:
                        continue
                e = rel(v, clones[k])
                if e > (2e-3 if (v.dtype == torch.float32 and v.dim() == 1) else 2e-4):
                    bad.append((name, k, e, tuple(v.shape)))
        return r
    emu.call = checked
    prev = _lib._backend
    _lib.set_backend(emu)
    try:
        torch.manual_seed(0)
        model = fastscnn(3, 19).train()
        opt = FlatAdamW(model.parameters(), lr=1e-3)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 3, 32, 64, generator=g)
        y = torch.randint(0, 19, (2, 32, 64), generator=g)
        y[0, :4] = 255
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        opt.zero_grad()
        loss = CrossEntropyLoss(ignore_index=255)(model(x), y)
        loss.backward()
        opt.step()
        assert torch.isfinite(loss.detach())
        cm = ConfusionMatrix(19)
        with torch.no_grad():
            cm.update((model.eval()(x), y))
        assert int(cm.compute(sync=False).sum()) == int((y != 255).sum())
    finally:
        _lib.set_backend(prev)
    assert not bad, bad[:5]
    assert len(seen) >= 20 and emu.emulated_calls > 330, (len(seen), emu.emulated_calls)
