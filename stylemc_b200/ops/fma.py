"""fma: a * b + c with a broadcast-aware backward.  Same surface as the reference ``torch_utils/ops/fma.py:15-58``.

On the fused synthesis path the multiply-add (demodulation coefficient times conv result plus noise) lives in the
epilogue of the implicit-GEMM kernel; this function is the stand-alone op for callers of the reference API.
"""
import torch


def fma(a, b, c):
    """fma.py:15-16."""
    return _FusedMultiplyAdd.apply(a, b, c)


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        out = torch.addcmul(c, a, b)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        da = db = dc = None
        if ctx.needs_input_grad[0]:
            da = _unbroadcast(dout * b, a.shape)
        if ctx.needs_input_grad[1]:
            db = _unbroadcast(dout * a, b.shape)
        if ctx.needs_input_grad[2]:
            dc = _unbroadcast(dout, ctx.c_shape)
        return da, db, dc


def _unbroadcast(x, shape):
    """Sum ``x`` over the axes that broadcasting expanded (fma.py:49-58)."""
    extra = x.ndim - len(shape)
    assert extra >= 0
    dims = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra or shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    assert x.shape == shape
    return x
