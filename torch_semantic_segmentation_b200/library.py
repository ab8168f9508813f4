"""``torch.library`` registration of the C ABI (namespace ``tss_b200``).

Two layers, both on top of the same ``libtss_b200.so`` that ``_lib`` loads with ctypes:

1. **Every kernel launcher of ``include/tss_b200.h``** becomes an out-variant operator
   ``torch.ops.tss_b200.<name without the tss_ prefix>(...) -> ()``: the schema is generated from the header (a ``const T*``
   parameter is an optional input tensor, a non-const pointer an optional MUTATED tensor ``Tensor(a!)?``, integers and
   floats are ``int`` / ``float``; ``stream`` does not appear -- the caller's current stream is used).  The
   implementation is the ctypes call; the fake (meta) implementation does nothing, which is exact for out-variant
   operators whose outputs the caller allocates.  ``_lib.call`` -- the only way the package reaches a kernel -- goes
   through these operators, so the model's forward and backward are sequences of registered ``tss_b200`` ops
   (visible to the dispatcher, to ``FakeTensorMode`` and to profilers by name).
2. **Functional, differentiable operators** with ``register_fake`` + ``register_autograd`` for the blocks a maintainer
   of the reference would call on their own (the reference's call sites in parentheses):
   ``dwconv3x3`` (fastscnn.py:179-180), ``pwconv`` (fastscnn.py:167-168, 97), ``bilinear`` (``F.interpolate(...,
   align_corners=True)``, fastscnn.py:74,119-120), ``upsample_logits`` (fastscnn.py:63-64), ``cross_entropy``
   (scripts/train_fastscnn.py:132).  The conv + BatchNorm blocks of the models stay ``torch.autograd.Function``s
   (functional.ConvBNAct): their context carries links between layers that are not tensors.
"""
import torch

from . import _lib

NS = 'tss_b200'
_def = torch.library.Library(NS, 'DEF')

# pointer parameters that are small HOST arrays read by the launcher itself (never a device pointer)
HOST_ARRAYS = {'bins': torch.int32, 'norm': torch.float32}


class _HostArray:
    """What the backends expect for a host-array argument: ``.values`` and a ctypes ``.array``."""

    def __init__(self, tensor):
        import ctypes
        self.values = tuple(tensor.tolist())
        ctype = ctypes.c_int if tensor.dtype == torch.int32 else ctypes.c_float
        self.array = (ctype * len(self.values))(*self.values)


def _schema(short, params):
    parts, alias = [], 0
    for pname, kind, base, const in params:
        if pname == 'stream':
            continue
        if kind == 'ptr':
            if const or pname in HOST_ARRAYS:
                parts.append('Tensor? %s' % pname)
            else:
                parts.append('Tensor(%s!)? %s' % (chr(ord('a') + alias), pname))
                alias += 1
        elif base in ('float', 'double'):
            parts.append('float %s' % pname)
        else:
            parts.append('int %s' % pname)
    return '%s(%s) -> ()' % (short, ', '.join(parts))


def _make_impl(name, names):
    def impl(*args):
        kwargs = {}
        for pname, v in zip(names, args):
            if pname in HOST_ARRAYS and isinstance(v, torch.Tensor):
                v = _HostArray(v)
            kwargs[pname] = v
        _lib.backend().call(name, kwargs)
    impl.__name__ = name
    return impl


def _nothing(*args):
    return None


OPS = {}        # C name -> (operator, parameter names in schema order)
_CAPTURE_CHECK = __import__('os').environ.get('TSS_CAPTURE_CHECK') == '1'


def _register_launchers():
    for name, (ret, params) in _lib.parse_header(with_const=True).items():
        if ret != 'int' or not any(p[0] == 'stream' for p in params):
            continue                      # version / error string / workspace size queries: plain ctypes calls
        short = name[4:]
        names = [p[0] for p in params if p[0] != 'stream']
        _def.define(_schema(short, params))
        _def.impl(short, _make_impl(name, names), 'CompositeExplicitAutograd')
        torch.library.register_fake('%s::%s' % (NS, short))(_nothing)
        OPS[name] = (getattr(getattr(torch.ops, NS), short).default, names)


_register_launchers()


def check_capture(where):
    """Debugging aid (TSS_CAPTURE_CHECK=1 runs it after every library call): raise as soon as the CUDA-graph capture of the
    current stream has been invalidated, so that the traceback names the first call after the offending one."""
    status = _lib.backend().call('tss_capture_status', {'stream_handle': torch.cuda.current_stream().cuda_stream})
    if status not in (0, 1):
        raise RuntimeError('CUDA graph capture invalidated at or before %s (status %d)' % (where, status))
    return status


def instrument_backward(root):
    """Debugging aid: ``check_capture`` before and after every node of the autograd graph under ``root``."""
    seen, stack = set(), [root]
    while stack:
        node = stack.pop()
        if node is None or node in seen:
            continue
        seen.add(node)
        name = type(node).__name__

        def pre(grads, name=name):
            check_capture('entering ' + name)

        def post(grad_inputs, grad_outputs, name=name):
            check_capture('leaving ' + name)
        node.register_prehook(pre)
        node.register_hook(post)
        stack.extend(nxt for nxt, _ in node.next_functions)
    return len(seen)


def dispatch(name, kwargs):
    """``_lib.call`` for a kernel launcher: positional call of the registered operator."""
    op, names = OPS[name]
    extra = set(kwargs) - set(names)
    if extra:
        raise TypeError('%s: unknown arguments %s' % (name, sorted(extra)))
    args = []
    for pname in names:
        if pname not in kwargs:
            raise TypeError('%s: missing argument %r' % (name, pname))
        v = kwargs[pname]
        if hasattr(v, 'array') and not isinstance(v, torch.Tensor):       # ops._HostInts / _HostFloats
            v = torch.tensor(v.values, dtype=HOST_ARRAYS[pname])
        args.append(v)
    op(*args)
    if _CAPTURE_CHECK:
        check_capture(name)
    return 0


# ------------------------------------------------------------------ functional operators ------------------------
def _ops():
    from . import ops
    return ops


_def.define('dwconv3x3(Tensor x, Tensor weight, int stride=1, int dilation=1) -> Tensor')
_def.define('pwconv(Tensor x, Tensor weight, Tensor? bias=None) -> Tensor')
_def.define('bilinear(Tensor x, int height, int width) -> Tensor')
_def.define('bilinear_backward(Tensor grad, int height, int width) -> Tensor')
_def.define('upsample_logits(Tensor scores, int height, int width) -> Tensor')
_def.define('upsample_logits_backward(Tensor grad, int height, int width, int pitch) -> Tensor')
_def.define('cross_entropy(Tensor logits, Tensor target, int ignore_index=-100) -> (Tensor, Tensor)')


def _nhwc_like(x, C, H, W, pitch=None):
    return _ops().empty_nhwc(x.shape[0], C, H, W, x.dtype, x.device, pitch)


def _out_hw(H, W, stride):
    return (H - 1) // stride + 1, (W - 1) // stride + 1


def _pw_pitch(Nc):
    return None if Nc % 8 == 0 else (Nc + 7) // 8 * 8 + 8


# ---- depthwise 3x3 (padding = dilation) -----------------------------------------------------------------------
@torch.library.impl(_def, 'dwconv3x3', 'CompositeExplicitAutograd')
def _dwconv3x3(x, weight, stride=1, dilation=1):
    ops = _ops()
    return ops.dwconv_fwd(ops.as_nhwc(x), weight, stride, dilation)


@torch.library.register_fake(NS + '::dwconv3x3')
def _dwconv3x3_fake(x, weight, stride=1, dilation=1):
    Ho, Wo = _out_hw(x.shape[2], x.shape[3], stride)
    return _nhwc_like(x, x.shape[1], Ho, Wo)


def _dwconv3x3_setup(ctx, inputs, output):
    x, weight, stride, dilation = inputs
    ctx.save_for_backward(x, weight)
    ctx.conf = (stride, dilation)


def _dwconv3x3_backward(ctx, grad):
    ops = _ops()
    x, weight = ctx.saved_tensors
    stride, dilation = ctx.conf
    grad = ops.as_nhwc(grad)
    dx = ops.dwconv_dgrad(grad, weight, x.shape[2], x.shape[3], stride, dilation) if ctx.needs_input_grad[0] else None
    dw = None
    if ctx.needs_input_grad[1]:
        dw = torch.zeros_like(weight)
        ops.dwconv_wgrad(ops.as_nhwc(x), grad, dw, stride, dilation)
    return dx, dw, None, None


torch.library.register_autograd(NS + '::dwconv3x3', _dwconv3x3_backward, setup_context=_dwconv3x3_setup)


# ---- pointwise 1x1 (+ bias) ------------------------------------------------------------------------------------
@torch.library.impl(_def, 'pwconv', 'CompositeExplicitAutograd')
def _pwconv(x, weight, bias=None):
    ops = _ops()
    return ops.pwconv_fwd(ops.as_nhwc(x), weight, shift=bias)


@torch.library.register_fake(NS + '::pwconv')
def _pwconv_fake(x, weight, bias=None):
    Nc = weight.shape[0]
    return _nhwc_like(x, Nc, x.shape[2], x.shape[3], _pw_pitch(Nc))


def _pwconv_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _pwconv_backward(ctx, grad):
    ops = _ops()
    x, weight = ctx.saved_tensors
    Nc = weight.shape[0]
    g = ops.geom(grad)
    if g is None or g[4] % 8 != 0 or g[4] < (Nc + 7) // 8 * 8:       # 16-byte aligned rows with zeroed pad columns
        buf = torch.zeros((grad.shape[0], grad.shape[2], grad.shape[3], (Nc + 7) // 8 * 8 + 8), dtype=grad.dtype, device=grad.device)
        view = buf[..., :Nc].permute(0, 3, 1, 2)
        view.copy_(grad)
        grad = view
    dx = ops.pwconv_dgrad(grad, weight) if ctx.needs_input_grad[0] else None
    dw = db = None
    if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
        dw = torch.zeros_like(weight)
        db = torch.zeros(Nc, dtype=torch.float32, device=weight.device) if ctx.has_bias else None
        ops.pwconv_wgrad(ops.as_nhwc(x), grad, dw, db)
    return dx, dw, db


torch.library.register_autograd(NS + '::pwconv', _pwconv_backward, setup_context=_pwconv_setup)


# ---- bilinear resize, align_corners=True ---------------------------------------------------------------------
@torch.library.impl(_def, 'bilinear', 'CompositeExplicitAutograd')
def _bilinear(x, height, width):
    ops = _ops()
    return ops.bilinear_fwd(ops.as_nhwc(x), height, width)


@torch.library.register_fake(NS + '::bilinear')
def _bilinear_fake(x, height, width):
    return _nhwc_like(x, x.shape[1], height, width)


@torch.library.impl(_def, 'bilinear_backward', 'CompositeExplicitAutograd')
def _bilinear_backward_impl(grad, height, width):
    ops = _ops()
    return ops.bilinear_bwd(ops.as_nhwc(grad), height, width)


@torch.library.register_fake(NS + '::bilinear_backward')
def _bilinear_backward_fake(grad, height, width):
    return _nhwc_like(grad, grad.shape[1], height, width)


def _bilinear_setup(ctx, inputs, output):
    ctx.in_hw = (inputs[0].shape[2], inputs[0].shape[3])


def _bilinear_backward(ctx, grad):
    return getattr(torch.ops, NS).bilinear_backward(grad, ctx.in_hw[0], ctx.in_hw[1]), None, None


torch.library.register_autograd(NS + '::bilinear', _bilinear_backward, setup_context=_bilinear_setup)


# ---- final up-sampling: NHWC class scores -> NCHW-contiguous logits ---------------------------------------------
@torch.library.impl(_def, 'upsample_logits', 'CompositeExplicitAutograd')
def _upsample_logits(scores, height, width):
    return _ops().upsample_logits_fwd(scores, height, width)


@torch.library.register_fake(NS + '::upsample_logits')
def _upsample_logits_fake(scores, height, width):
    return torch.empty((scores.shape[0], scores.shape[1], height, width), dtype=scores.dtype, device=scores.device)


@torch.library.impl(_def, 'upsample_logits_backward', 'CompositeExplicitAutograd')
def _upsample_logits_backward_impl(grad, height, width, pitch):
    return _ops().upsample_logits_bwd(grad, height, width, pitch)


@torch.library.register_fake(NS + '::upsample_logits_backward')
def _upsample_logits_backward_fake(grad, height, width, pitch):
    return _nhwc_like(grad, grad.shape[1], height, width, pitch)


def _upsample_logits_setup(ctx, inputs, output):
    scores = inputs[0]
    g = _ops().geom(scores)
    ctx.in_hw = (scores.shape[2], scores.shape[3])
    ctx.pitch = max(g[4] if g is not None else 0, (scores.shape[1] + 7) // 8 * 8)


def _upsample_logits_backward(ctx, grad):
    return getattr(torch.ops, NS).upsample_logits_backward(grad, ctx.in_hw[0], ctx.in_hw[1], ctx.pitch), None, None


torch.library.register_autograd(NS + '::upsample_logits', _upsample_logits_backward, setup_context=_upsample_logits_setup)


# ---- softmax cross-entropy with ignore_index: loss and its gradient in ONE pass --------------------------------
@torch.library.impl(_def, 'cross_entropy', 'CompositeExplicitAutograd')
def _cross_entropy(logits, target, ignore_index=-100):
    loss, dlogits, _, _ = _ops().ce_forward(logits, target, ignore_index, want_grad=True)
    return loss, dlogits


@torch.library.register_fake(NS + '::cross_entropy')
def _cross_entropy_fake(logits, target, ignore_index=-100):
    return (torch.empty((), dtype=torch.float32, device=logits.device),
            torch.empty(logits.shape, dtype=logits.dtype, device=logits.device))


def _cross_entropy_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.mark_non_differentiable(output[1])


def _cross_entropy_backward(ctx, grad_loss, _grad_dlogits):
    (dlogits,) = ctx.saved_tensors
    return dlogits * grad_loss.to(dlogits.dtype), None, None


torch.library.register_autograd(NS + '::cross_entropy', _cross_entropy_backward, setup_context=_cross_entropy_setup)
