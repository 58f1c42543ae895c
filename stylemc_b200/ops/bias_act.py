"""bias_act: fused bias + activation + gain + clamp, CUDA only.

Same call surface as the reference ``torch_utils/ops/bias_act.py:55-89`` (``bias_act(x, b, dim, act, alpha,
gain, clamp, impl)``) and the same ``activation_funcs`` table (:23-33), backed by ``smc_bias_act``.
First- and second-order gradients follow the reference's Function pair (:129-210).  ``impl='ref'`` is not
provided: the package has no non-CUDA implementation (ask the oracle in tests instead).
"""
import math
from types import SimpleNamespace

import torch

from .. import _lib

activation_funcs = {
    'linear':   SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     SimpleNamespace(def_alpha=0,   def_gain=math.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    SimpleNamespace(def_alpha=0.2, def_gain=math.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    SimpleNamespace(def_alpha=0,   def_gain=math.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}


def _launch(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
    """One kernel pass; mirrors `_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)`."""
    _lib.require_cuda(x, 'x')
    if x.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f'bias_act: unsupported dtype {x.dtype}')
    if x.numel() > 2 ** 31 - 1:
        raise RuntimeError('x is too large')                                   # bias_act.cpp:40
    for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
        if t is not None and (t.shape != x.shape or t.dtype != x.dtype or t.device != x.device or t.stride() != x.stride()):
            raise RuntimeError(f'{name} must have the same shape, dtype, device and layout as x')   # bias_act.cpp:37-39,46-51
    if b is not None:
        if b.ndim != 1:
            raise RuntimeError('b must have rank 1')                           # bias_act.cpp:41
        if b.dtype != x.dtype or b.device != x.device:
            raise RuntimeError('b must have the same dtype and device as x')   # bias_act.cpp:36
        if not 0 <= dim < x.ndim:
            raise RuntimeError('dim is out of bounds')                         # bias_act.cpp:42
        if b.numel() != x.shape[dim]:
            raise RuntimeError('b has wrong number of elements')               # bias_act.cpp:43
        b = b.contiguous()
    y = torch.empty_like(x)
    step_b = x.stride(dim) if b is not None else 1
    with torch.cuda.device(x.device):
        _lib.call('smc_bias_act', _lib.ptr(x), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), _lib.ptr(y),
                  _lib.DTYPE_CODE[x.dtype], x.numel(), b.numel() if b is not None else 0, step_b, grad, spec.cuda_idx,
                  alpha, gain, clamp, _lib.stream())
    return y


def _dense(t):
    """Keep channels_last if that is what we got, else make contiguous (bias_act.py:148-149)."""
    if t.ndim > 2 and t.stride(1) == 1:
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


_cache = {}


def _functions(dim, act, alpha, gain, clamp):
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]
    spec = activation_funcs[act]
    needs_x = 'x' in spec.ref or spec.has_2nd_grad
    # 'linear' needs nothing for its slope, but the clamp mask is defined on y (the reference's impl='ref' path gets it from
    # autograd of clamp(); its plugin skips it) -- parity is stated against impl='ref', so keep y whenever a clamp is active.
    needs_y = 'y' in spec.ref or (spec.ref == '' and clamp >= 0)

    class BiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            x = _dense(x)
            y = x
            if act != 'linear' or gain != 1 or clamp >= 0 or b is not None:
                y = _launch(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(x if needs_x else None, b if needs_x else None, y if needs_y else None)
            ctx.has_b = b is not None
            return y

        @staticmethod
        def backward(ctx, dy):
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or (ctx.has_b and ctx.needs_input_grad[1]):
                dx = dy
                if act != 'linear' or gain != 1 or clamp >= 0:
                    dx = BiasActGrad.apply(dy, x, b, y)
            if ctx.has_b and ctx.needs_input_grad[1]:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ref = y if y is not None else x
            if ref is None:      # 'linear': the gradient depends on neither x nor y (bias_act.py:26), keep dy's own layout
                dy = _dense(dy)
            else:
                dy = dy.contiguous(memory_format=torch.channels_last) if (ref.ndim > 2 and ref.stride(1) == 1) else dy.contiguous()
            ctx.memory_format = torch.channels_last if (dy.ndim > 2 and dy.stride(1) == 1) else torch.contiguous_format   # bias_act.py:171
            dx = _launch(dy, b, x, y, None, 1, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            dy, x, b, y = ctx.saved_tensors
            d_dx = d_dx.contiguous(memory_format=ctx.memory_format)          # bias_act.py:184: the layout of the saved x / y
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _launch(d_dx, b, x, y, dy, 2, dim, spec, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _cache[key] = BiasAct
    return BiasAct


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain, +-clamp).  Arguments as bias_act.py:55-84; CUDA tensors only."""
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda':
        raise RuntimeError("stylemc_b200 only implements impl='cuda' (no reference/CPU path in the product)")
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(spec.def_alpha if alpha is None else alpha)
    gain = float(spec.def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    return _functions(dim, act, alpha, gain, clamp).apply(x, b)
