"""The C-ABI library: it loads, exports every symbol the header declares, and refuses to run
without a GPU instead of falling back."""
import ctypes

import pytest
import torch

from torch_semantic_segmentation_b200 import _lib, ops


def test_header_parses_and_library_exports_every_symbol(built_lib):
    protos = _lib.parse_header()
    assert len(protos) >= 35
    lib = ctypes.CDLL(built_lib)
    for name in protos:
        assert hasattr(lib, name), name
    lib.tss_version.restype = ctypes.c_int
    assert lib.tss_version() == 100
    lib.tss_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.tss_last_error(), bytes)


def test_argument_errors_are_reported_through_the_abi(built_lib):
    lib = ctypes.CDLL(built_lib)
    lib.tss_last_error.restype = ctypes.c_char_p
    # C = 12 is not a multiple of 8 -> hard error, message available (no kernel is launched)
    rc = lib.tss_dwconv3x3_fwd(None, None, None, 1, 8, 8, 12, 1, 1, None, None, 0, None, 0, None)
    assert rc == 1 and b'multiple of 8' in lib.tss_last_error()
    rc = lib.tss_dwconv3x3_fwd(None, None, None, 1, 8, 8, 16, 3, 1, None, None, 0, None, 0, None)
    assert rc == 1 and b'unsupported stride' in lib.tss_last_error()
    rc = lib.tss_bn_finalize(None, ctypes.c_int64(1), None, None, None, None, None, ctypes.c_float(0.1),
                             ctypes.c_float(1e-5), None, None, None, None, 8, None)
    assert rc == 1 and b'more than 1 value per channel' in lib.tss_last_error()


def test_no_cpu_path(built_lib):
    prev = _lib._backend
    _lib.set_backend(None)
    try:
        x = ops.empty_nhwc(1, 8, 4, 4, torch.float32, 'cpu')
        w = torch.zeros(8, 1, 3, 3)
        with pytest.raises(RuntimeError, match='CUDA tensor'):
            ops.dwconv_fwd(x, w, 1, 1)
    finally:
        _lib.set_backend(prev)


def test_named_argument_checking(built_lib):
    b = _lib._Backend()
    with pytest.raises(TypeError, match='missing argument'):
        b.call('tss_bn_fold', dict(gamma=None))
    with pytest.raises(TypeError, match='unknown arguments'):
        b.call('tss_ce_finalize', dict(loss_sum=None, nvalid=None, loss=None, bogus=1))
