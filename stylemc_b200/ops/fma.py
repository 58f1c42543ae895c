"""fma: ``a * b + c`` with broadcasting, forward and backward on repo kernels (``smc_fma`` / ``smc_fma_reduce``, csrc/misc.cu).

Call surface of the reference ``torch_utils/ops/fma.py`` (``fma(a, b, c)``, :15).  There the forward is ``torch.addcmul`` and the
backward multiplies in full size and then sums the broadcast axes away (:36-58).  Here the forward is one strided kernel over the
broadcast index space (operands are never materialised at full size) and each input gradient is ONE multiply-reduce kernel that sums
``dout * other`` straight into the operand's own shape.  On the fused synthesis path this multiply-add (demodulation coefficient x
conv result + noise) is the epilogue of the implicit-GEMM kernel; this module serves callers of the op-level API (networks.py).
"""
import ctypes

import torch

from .. import _lib

_RANK = 4
_I64x4 = ctypes.c_int64 * _RANK


def _space(*tensors):
    """Common broadcast shape, left-padded to rank 4."""
    shape = torch.broadcast_shapes(*[t.shape for t in tensors])
    if len(shape) > _RANK:
        raise RuntimeError(f'fma: at most {_RANK} dimensions are supported, got {len(shape)}')
    return (1,) * (_RANK - len(shape)) + tuple(shape)


def _strides(t, space):
    """Element strides of ``t`` viewed in ``space``: 0 on the axes it is broadcast along (and on size-1 axes)."""
    view = t.reshape((1,) * (_RANK - t.ndim) + tuple(t.shape))
    return [0 if (view.shape[k] == 1) else view.stride(k) for k in range(_RANK)]


def _check(*tensors):
    ref = tensors[0]
    for t in tensors:
        _lib.require_cuda(t, 'fma operand')
        if t.dtype != ref.dtype or t.device != ref.device:
            raise RuntimeError('fma: operands must share dtype and device')
    if ref.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f'fma: unsupported dtype {ref.dtype}')


def _forward(a, b, c):
    _check(a, b, c)
    space = _space(a, b, c)
    out = torch.empty(torch.broadcast_shapes(a.shape, b.shape, c.shape), dtype=a.dtype, device=a.device)
    if out.numel() == 0:
        return out
    # the four host arrays must outlive the call: keep them in locals (ctypes.addressof of a temporary dangles)
    shape, sa, sb, sc = _I64x4(*space), _I64x4(*_strides(a, space)), _I64x4(*_strides(b, space)), _I64x4(*_strides(c, space))
    with torch.cuda.device(a.device):
        _lib.call('smc_fma', _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(out), _lib.DTYPE_CODE[a.dtype],
                  ctypes.addressof(shape), ctypes.addressof(sa), ctypes.addressof(sb), ctypes.addressof(sc), _lib.stream())
    return out


def _reduce_to(x, y, shape):
    """sum over the broadcast axes of ``x * y`` (``y`` may be None) into a new tensor of ``shape`` (the role of fma.py:49-58)."""
    _check(*([x] if y is None else [x, y]))
    space = _space(x) if y is None else _space(x, y)
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    if out.numel() == 0:
        return out
    if x.numel() == 0:
        return out.zero_()
    out_view = out.reshape((1,) * (_RANK - out.ndim) + tuple(out.shape))
    so = [0 if (out_view.shape[k] == 1 and space[k] > 1) else out_view.stride(k) for k in range(_RANK)]
    shape, sx, sout = _I64x4(*space), _I64x4(*_strides(x, space)), _I64x4(*so)          # kept alive across the call
    sy = None if y is None else _I64x4(*_strides(y, space))
    with torch.cuda.device(x.device):
        _lib.call('smc_fma_reduce', _lib.ptr(x), _lib.ptr(y), _lib.ptr(out), _lib.DTYPE_CODE[x.dtype], ctypes.addressof(shape),
                  ctypes.addressof(sx), None if sy is None else ctypes.addressof(sy), ctypes.addressof(sout), _lib.stream())
    return out


class _Fma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return _forward(a, b, c)

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        need_a, need_b, need_c = ctx.needs_input_grad
        return (_FmaReduce.apply(grad_out, b, a.shape) if need_a else None,      # d/da = sum_bcast(dout * b), fma.py:36-37
                _FmaReduce.apply(grad_out, a, b.shape) if need_b else None,      # d/db = sum_bcast(dout * a), fma.py:39-40
                _FmaReduce.apply(grad_out, None, ctx.c_shape) if need_c else None)   # d/dc = sum_bcast(dout), fma.py:42-43


class _FmaReduce(torch.autograd.Function):
    """out = sum_bcast(x * y) -> ``shape``; differentiable itself (second-order terms are again fma kernels)."""

    @staticmethod
    def forward(ctx, x, y, shape):
        ctx.save_for_backward(x, y)
        return _reduce_to(x, y, shape)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        g = g.contiguous()
        zero = torch.zeros([], dtype=g.dtype, device=g.device)
        gx = gy = None
        if ctx.needs_input_grad[0]:      # d/dx = g (broadcast back) * y
            one = torch.ones([], dtype=g.dtype, device=g.device)
            gx = _Fma.apply(g.expand(x.shape) if g.shape != x.shape else g, y.expand(x.shape) if y is not None else one, zero)
        if y is not None and ctx.needs_input_grad[1]:
            gy = _FmaReduce.apply(_Fma.apply(g, x, zero), None, y.shape)
        return gx, gy, None


def fma(a, b, c):
    """``a * b + c`` (fma.py:15)."""
    return _Fma.apply(a, b, c)
