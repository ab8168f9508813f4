"""ctypes binding of ``libtss_b200.so`` (the C-ABI CUDA library, ``include/tss_b200.h``).

The prototypes are parsed from the header, so the Python side can never drift from the
C ABI: every call is made with *named* arguments that are matched against the header's
parameter names and checked against its C types (a ``const float*`` only accepts a
float32 tensor, an ``int64_t*`` an int64 tensor, ...).  Tensors are passed as raw device
pointers; ``stream`` is filled in with the caller's current CUDA stream.

There is no fallback: if the library is missing, or a call returns non-zero, a
``RuntimeError`` carrying ``tss_last_error()`` is raised.
"""
import ctypes
import os
import re
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TSS_LIB') or os.path.join(_HERE, 'libtss_b200.so')      # TSS_LIB: tools/trace_kernels.py
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'tss_b200.h')

TSS_F32, TSS_BF16 = 0, 1
EPI_RELU = 1

_DTYPE_CODE = {torch.float32: TSS_F32, torch.bfloat16: TSS_BF16}

_CTYPES = {
    'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64,
    'float': ctypes.c_float, 'double': ctypes.c_double, 'size_t': ctypes.c_size_t,
}
# pointer C type -> torch dtype it must point at (None = any)
_PTR_DTYPE = {
    'void': None, 'float': torch.float32, 'double': torch.float64,
    'int64_t': torch.int64, 'int': torch.int32,
}

_PROTO_RE = re.compile(r'^\s*(int|int64_t|uint64_t|const char\*)\s+(tss_\w+)\s*\(([^)]*)\)\s*;', re.M | re.S)


def dtype_code(dtype):
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise RuntimeError('tss_b200: unsupported activation dtype %s (float32 or bfloat16)' % dtype)


def parse_header(path=HEADER_PATH, with_const=False):
    """-> {name: (restype, [(param_name, kind, base_type)])}, kind in {'ptr', 'val'}; ``with_const`` appends whether
    the parameter is declared ``const`` (an input) -- library.py derives the operators' mutation annotations from it."""
    with open(path) as f:
        text = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    protos = {}
    for ret, name, args in _PROTO_RE.findall(text):
        params = []
        args = ' '.join(args.split())
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                m = re.match(r'^(const\s+)?(\w+)\s*(\*?)\s*(\w+)$', a)
                if not m:
                    raise RuntimeError('cannot parse parameter %r of %s' % (a, name))
                const, base, star, pname = m.groups()
                params.append((pname, 'ptr' if star else 'val', base) + ((bool(const),) if with_const else ()))
        protos[name] = (ret, params)
    return protos


class _Backend:
    """Real backend: raw pointers into libtss_b200.so."""

    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'tss_b200: %s is missing -- build it with `make` (or '
                '`python -c "import __graft_entry__ as g; g.build()"`); there is no fallback path'
                % LIB_PATH)
        self.lib = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.fns = {}
        for name, (ret, params) in self.protos.items():
            fn = getattr(self.lib, name)      # AttributeError = header/library drift
            fn.restype = {'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64,
                          'const char*': ctypes.c_char_p}[ret]
            fn.argtypes = [ctypes.c_void_p if kind == 'ptr' else _CTYPES[base]
                           for _, kind, base in params]
            self.fns[name] = fn
        if self.lib.tss_version() != 100:
            raise RuntimeError('tss_b200: library/header version mismatch')

    def call(self, name, kwargs):
        ret, params = self.protos[name]
        args = []
        device = None
        for pname, kind, base in params:
            if pname == 'stream':
                continue
            if pname not in kwargs:
                raise TypeError('%s: missing argument %r' % (name, pname))
            v = kwargs[pname]
            if kind == 'ptr':
                if v is None:
                    args.append(None)
                    continue
                if hasattr(v, 'array'):          # small host array read by the launcher itself
                    args.append(ctypes.cast(v.array, ctypes.c_void_p))
                    continue
                if not isinstance(v, torch.Tensor):
                    raise TypeError('%s: %s must be a tensor or None' % (name, pname))
                if not v.is_cuda:
                    raise RuntimeError('%s: %s must be a CUDA tensor (no CPU path exists)' % (name, pname))
                want = _PTR_DTYPE[base]
                if want is not None and v.dtype != want:
                    raise TypeError('%s: %s must be %s, got %s' % (name, pname, want, v.dtype))
                device = v.device if device is None else device
                if v.device != device:
                    raise RuntimeError('%s: tensors on different devices' % name)
                args.append(v.data_ptr())
            else:
                args.append(v)
        extra = set(kwargs) - {p[0] for p in params}
        if extra:
            raise TypeError('%s: unknown arguments %s' % (name, sorted(extra)))
        if any(p[0] == 'stream' for p in params):
            with torch.cuda.device(device):
                args.append(torch.cuda.current_stream(device).cuda_stream)
                rc = self.fns[name](*args)
        else:
            rc = self.fns[name](*args)
        if ret == 'int' and rc != 0:
            raise RuntimeError('%s failed (%d): %s' % (name, rc, self.lib.tss_last_error().decode()))
        return rc


_backend = None
_lock = threading.Lock()


def backend():
    global _backend
    if _backend is None:
        with _lock:
            if _backend is None:
                _backend = _Backend()
    return _backend


def set_backend(b):
    """Test hook (tests/fake_backend.py): install an object with ``call(name, kwargs)``."""
    global _backend
    _backend = b


def call(name, **kwargs):
    """Run one entry point of the C ABI.  Kernel launchers go through their registered ``torch.ops.tss_b200`` operator
    (library.py: schema + fake implementation generated from the header), whose implementation is the ctypes call."""
    from . import library
    if name in library.OPS:
        return library.dispatch(name, kwargs)
    return backend().call(name, kwargs)


def launch_count():
    b = backend()
    return int(b.lib.tss_launch_count()) if hasattr(b, 'lib') else 0
