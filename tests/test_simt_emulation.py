"""The plain SIMT kernels that have not run on a B200 yet (csrc/ppm.cu, csrc/augment.cu), executed on the CPU by
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
SOURCES = ['ppm.cu', 'augment.cu']


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


class EmulatedBackend:
    """Entry points exported by the host build run the real kernel code; everything else is the torch emulation."""

    def __init__(self, so_path):
        self.lib = ctypes.CDLL(so_path)
        self.protos = _lib.parse_header()
        self.fake = FakeBackend()
        self.emulated_calls = 0

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


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(3, 6, 9), (2, 8, 8), (20, 1, 2)])
def test_grouped_pyramid_kernels_on_the_simt_emulation(emulated, shape, dtype):
    N, H, W = shape
    before = emulated.emulated_calls
    got = _ppm_run(emulated, True, N, H, W, dtype)
    assert emulated.emulated_calls - before == 4          # branches fwd, concat fwd, concat bwd, branches bwd
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
