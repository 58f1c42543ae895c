"""upfirdn2d: pad -> zero-upsample -> FIR -> decimate, CUDA only.

Same call surface as the reference ``torch_utils/ops/upfirdn2d.py`` (``setup_filter`` :72, ``upfirdn2d`` :120,
``filter2d`` :272, ``upsample2d`` :308, ``downsample2d`` :347), backed by ``smc_upfirdn2d``.  The backward is,
as in the reference (:245-264), another upfirdn2d with up/down swapped, the filter flipped and padding
recomputed, so gradients of any order work.
"""
import numpy as np
import torch

from .. import _lib


def _scaling(v):
    if isinstance(v, int):
        v = [v, v]
    assert isinstance(v, (list, tuple)) and all(isinstance(q, int) for q in v)
    sx, sy = v
    assert sx >= 1 and sy >= 1
    return sx, sy


def _padding(p):
    if isinstance(p, int):
        p = [p, p]
    assert isinstance(p, (list, tuple)) and all(isinstance(q, int) for q in p)
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    px0, px1, py0, py1 = p
    return px0, px1, py0, py1


def _filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR taps -> float32 filter tensor (2-D non-separable, or 1-D when separable); upfirdn2d.py:72-116."""
    f = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    assert f.ndim in (0, 1, 2) and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


_sep_cache = {}


def _separable_factors(f):
    """(fy[0..3], fx[0..3]) with f[i][j] == fy[i] * fx[j] for a rank-1 4x4 filter (setup_filter([1,3,3,1]) is one), else None.
    The filter lives on the device; the answer is cached per filter tensor so the host copy happens once.  The cache entry keeps
    the filter tensor alive: its address can then not be handed to another filter by the caching allocator (a freed 4x4 filter's
    address is what the next 4x4 filter gets, with ``_version`` 0 again), and an in-place change bumps ``_version``."""
    if tuple(f.shape) != (4, 4):
        return None
    key = (f.data_ptr(), f._version, f.device)
    if key not in _sep_cache:
        fh = f.detach().double().cpu()
        tot = float(fh.sum())
        res = None
        if abs(tot) > 1e-20:
            fy, fx = fh.sum(1) / tot, fh.sum(0)
            if float((torch.outer(fy, fx) - fh).abs().max()) <= 1e-6 * float(fh.abs().max()):
                res = [float(v) for v in fy] + [float(v) for v in fx]
        if len(_sep_cache) > 64:
            _sep_cache.clear()
        _sep_cache[key] = (res, f)
    return _sep_cache[key][0]


def _launch(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    """One pass; mirrors `_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)`."""
    _lib.require_cuda(x, 'x')
    if f.device != x.device:
        raise RuntimeError('f must reside on the same device as x')            # upfirdn2d.cpp:20
    if f.dtype != torch.float32:
        raise RuntimeError('f must be float32')                                # upfirdn2d.cpp:21
    if x.ndim != 4:
        raise RuntimeError('x must be rank 4')                                 # upfirdn2d.cpp:24
    if f.ndim != 2:
        raise RuntimeError('f must be rank 2')                                 # upfirdn2d.cpp:25
    if x.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f'upfirdn2d: unsupported dtype {x.dtype}')
    if x.numel() > 2 ** 31 - 1:
        raise RuntimeError('x is too large')                                   # upfirdn2d.cpp:22
    n, c, ih, iw = x.shape
    ow = (iw * upx + padx0 + padx1 - f.shape[1] + downx) // downx              # upfirdn2d.cpp:32-33
    oh = (ih * upy + pady0 + pady1 - f.shape[0] + downy) // downy
    if ow < 1 or oh < 1:
        raise RuntimeError('output must be at least 1x1')                      # upfirdn2d.cpp:34
    fmt = torch.channels_last if (x.stride(1) == 1 and c > 1) else torch.contiguous_format
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device, memory_format=fmt)
    p = _lib.UpfirdnParams()
    p.N, p.C, p.inH, p.inW, p.outH, p.outW = n, c, ih, iw, oh, ow
    p.x_stride[:] = list(x.stride())
    p.y_stride[:] = list(y.stride())
    p.fH, p.fW = f.shape
    p.f_stride[:] = list(f.stride())
    p.upx, p.upy, p.downx, p.downy, p.padx0, p.pady0, p.flip, p.gain = upx, upy, downx, downy, padx0, pady0, int(bool(flip)), float(gain)
    sep = _separable_factors(f)
    if sep is not None:
        p.separable = 1
        p.fsep[:] = sep
    import ctypes
    with torch.cuda.device(x.device):
        _lib.call('smc_upfirdn2d', _lib.ptr(x), _lib.ptr(f), _lib.ptr(y), _lib.DTYPE_CODE[x.dtype], ctypes.addressof(p), _lib.stream())
    return y


_cache = {}


def _function(up, down, padding, flip_filter, gain):
    upx, upy = _scaling(up)
    downx, downy = _scaling(down)
    padx0, padx1, pady0, pady1 = _padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _cache:
        return _cache[key]

    class Upfirdn2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
            if f.ndim == 2:
                y = _launch(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:  # separable: a row pass then a column pass, sqrt(gain) each (upfirdn2d.py:239-240)
                y = _launch(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, np.sqrt(gain))
                y = _launch(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, np.sqrt(gain))
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1,
                 fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]       # upfirdn2d.py:251-256
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _function(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter), gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, None

    _cache[key] = Upfirdn2d
    return Upfirdn2d


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Arguments as upfirdn2d.py:120-159; CUDA tensors only."""
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda':
        raise RuntimeError("stylemc_b200 only implements impl='cuda' (no reference/CPU path in the product)")
    return _function(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    upx, upy = _scaling(up)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    downx, downy = _scaling(down)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
