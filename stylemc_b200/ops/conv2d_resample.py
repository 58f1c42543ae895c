"""conv2d_resample: 2-D convolution with optional up/down-sampling, CUDA only.

Same call surface and dispatch as the reference ``torch_utils/ops/conv2d_resample.py:59-154``
(``conv2d_resample(x, w, f, up, down, padding, groups, flip_weight, flip_filter)``); the convolutions the
reference hands to cuDNN through ``_conv2d_wrapper`` (:29-54) run on the tcgen05 implicit-GEMM kernel
(``smc_igemm``) and the FIR passes on ``smc_upfirdn2d``.  NCHW in, NCHW out, dtype in = dtype out.  The GEMM
operands are fp16 hi+lo split planes (``precision='x3'``: three MMAs, ~21 mantissa bits), so fp32 callers get
fp32-grade results.  Gradient w.r.t. the input is implemented (same kernel, transposed weights); the path this
package accelerates keeps G frozen, so asking for a weight gradient raises.
"""
import torch

from .. import _lib, gemm
from . import conv2d_gradfix, upfirdn2d
from .upfirdn2d import _filter_size, _padding

PRECISION = 'x3'   # operand precision of the op-level API ('x1' = fp16 operands, 'x3' = split fp16)


def _ceil32(v):
    return (v + 31) // 32 * 32


def _pack(x, planes):
    """NCHW float tensor -> stacked NHWC fp16 planes [P*N, H, W, Cp] (Cp = channels rounded up to 32, zero filled)."""
    n, c, h, w = x.shape
    cp = _ceil32(c)
    x = x.float().contiguous()
    out = torch.zeros([planes * n, h, w, cp], dtype=torch.float16, device=x.device) if cp != c else \
        torch.empty([planes * n, h, w, cp], dtype=torch.float16, device=x.device)
    lo = out[n:] if planes == 2 else None
    with torch.cuda.device(x.device):
        _lib.call('smc_pack_nhwc', _lib.ptr(x), c * h * w, None, 0, _lib.ptr(out), _lib.ptr(lo), n, c, h * w, cp, _lib.stream())
    return out


def _unpack(y_nhwc, c, dtype):
    """NHWC fp32 [N, H, W, Cp] -> NCHW [N, c, H, W]."""
    n, h, w, cp = y_nhwc.shape
    out = torch.empty([n, c, h, w], dtype=torch.float32, device=y_nhwc.device)
    with torch.cuda.device(out.device):
        _lib.call('smc_unpack_nchw', _lib.ptr(y_nhwc), 0, _lib.ptr(out), None, n, c, h * w, cp, _lib.stream())
    return out.to(dtype)


def _weight_matrix(w, planes, transpose_io=False):
    """[O, I, kh, kw] -> K-major tap matrix planes [P * kh*kw*Op, Ip]: row (t*Op + o), col i.  With transpose_io the roles
    of O and I are swapped (dgrad: contraction over the output channels).  One ``smc_prepare_weights`` launch."""
    Bf, Bb, _, _ = gemm.prepare_weights(w.float(), two=planes == 2, fwd=not transpose_io, bwd=transpose_io, pad_to=32)
    return (Bb, _ceil32(w.shape[1])) if transpose_io else (Bf, _ceil32(w.shape[0]))


def _conv_fwd_raw(x, w, padding, transpose):
    """groups == 1 correlation (F.conv2d semantics) or stride-2 transposed conv (F.conv_transpose2d semantics,
    w = [I, O, kh, kw])."""
    planes = 2 if PRECISION == 'x3' else 1
    n, c, h, wd = x.shape
    A = _pack(x, planes)
    if not transpose:
        o, i, kh, kw = w.shape
        assert i == c
        py, px = padding
        ho, wo = h + 2 * py - kh + 1, wd + 2 * px - kw + 1
        if ho < 1 or wo < 1:
            raise RuntimeError('conv2d: output must be at least 1x1')
        B, op = _weight_matrix(w, planes)
        taps = [(ky - py, kx - px, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
        y = torch.empty([n, ho, wo, op], dtype=torch.float32, device=x.device)
        gemm.igemm(A, B, n, ho, wo, op, taps, precision=PRECISION, out_f32=y)
        return _unpack(y, o, x.dtype)
    i, o, kh, kw = w.shape
    assert i == c and kh == 3 and kw == 3, 'transposed conv: 3x3 stride 2 only'
    B, op = _weight_matrix(w, planes, transpose_io=True)
    ho, wo = 2 * h + 1, 2 * wd + 1
    y = torch.empty([n, ho, wo, op], dtype=torch.float32, device=x.device)
    for r in (0, 1):
        for cc in (0, 1):
            gemm.igemm(A, B, n, h + 1 - r, wd + 1 - cc, op, gemm.up2_parity_taps(r, cc), precision=PRECISION, out_f32=y,
                       out_strides=(ho * wo * op, 2 * wo * op, 2 * op), out_offset=(r * wo + cc) * op)
    return _unpack(y, o, x.dtype)


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, padding, transpose):
        ctx.save_for_backward(w)
        ctx.padding, ctx.transpose, ctx.x_shape = padding, transpose, x.shape
        return _conv_fwd_raw(x, w, padding, transpose)

    @staticmethod
    def backward(ctx, gy):
        w, = ctx.saved_tensors
        if ctx.needs_input_grad[1] and not conv2d_gradfix.weight_gradients_disabled:
            raise RuntimeError('stylemc_b200 conv2d: weight gradients are not implemented (the S-space path keeps G frozen); '
                               'wrap the call in conv2d_gradfix.no_weight_gradients()')
        gx = None
        if ctx.needs_input_grad[0]:
            if not ctx.transpose:
                kh, kw = w.shape[2:]
                py, px = ctx.padding
                gx = _Conv2d.apply(gy, w.transpose(0, 1).flip([2, 3]), (kh - 1 - py, kw - 1 - px), False)
            else:
                # gx[i, j] = sum_{ky,kx} gy[2i+ky, 2j+kx] * w[ky, kx]: a 3x3 correlation sampled on the even grid
                full = _Conv2d.apply(gy, w, (0, 0), False)
                gx = full[:, :, ::2, ::2]
        return gx, None, None, None


def _conv2d(x, w, padding=(0, 0), groups=1, transpose=False):
    _lib.require_cuda(x, 'x')
    if w.device != x.device:
        raise RuntimeError('w must reside on the same device as x')
    if groups == 1:
        return _Conv2d.apply(x, w, tuple(padding), transpose)
    # grouped: one launch per group (the fused modulated conv never comes here; see networks.modulated_conv2d)
    xs = x.chunk(groups, dim=1)
    ws = w.chunk(groups, dim=0)
    return torch.cat([_Conv2d.apply(a.contiguous(), b, tuple(padding), transpose) for a, b in zip(xs, ws)], dim=1)


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:29-54."""
    if not flip_weight:
        w = w.flip([2, 3])
    padding = (padding, padding) if isinstance(padding, int) else tuple(padding)
    if transpose:
        if stride != 2 or padding != (0, 0):
            raise RuntimeError('stylemc_b200: transposed conv supports stride 2, padding 0')
        return _conv2d(x, w, groups=groups, transpose=True)
    if stride != 1:
        raise RuntimeError('stylemc_b200: strided (down-sampling) convolution is not on the S-space path and not implemented')
    return _conv2d(x, w, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """Arguments as conv2d_resample.py:59-83."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    cout, cin_g, kh, kw = w.shape
    fw, fh = _filter_size(f)
    px0, px1, py0, py1 = _padding(padding)
    if up > 1:   # :84-88
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:  # :89-93
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2

    if kw == 1 and kh == 1 and down > 1 and up == 1:  # :96-99
        x = upfirdn2d.upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:  # :102-105
        x = _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:  # :108-111
        x = upfirdn2d.upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:  # :114-133
        if groups == 1:
            wt = w.transpose(0, 1)
        else:
            wt = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2).reshape(groups * cin_g, cout // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x, wt, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:  # :136-138
        return _conv2d_wrapper(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d.upfirdn2d(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)  # :141-145
    x = _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d.upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x
