"""Oracle: pad -> zero-upsample -> 2-D FIR -> decimate  (TEST INFRASTRUCTURE ONLY).

Restates the reference's slow path ``torch_utils/ops/upfirdn2d.py:168-208`` (``_upfirdn2d_ref``)
and the helpers ``setup_filter`` (:72-116), ``filter2d`` (:272-304), ``upsample2d`` (:308-343),
``downsample2d`` (:347-382); argument parsing rules from :37-68.
"""
import torch
import torch.nn.functional as F


def _pair(v):
    # upfirdn2d.py:37-44 (_parse_scaling): int -> (x, y)
    if isinstance(v, int):
        return v, v
    sx, sy = v
    return int(sx), int(sy)


def _pad4(p):
    # upfirdn2d.py:46-56 (_parse_padding): int | [x, y] | [x0, x1, y0, y1]
    if isinstance(p, int):
        return p, p, p, p
    p = [int(q) for q in p]
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    return tuple(p)


def filter_size(f):
    # upfirdn2d.py:58-68: returns (fw, fh)
    if f is None:
        return 1, 1
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(taps, normalize=True, flip_filter=False, gain=1.0, separable=None, dtype=torch.float32):
    """upfirdn2d.py:72-116.  [1,3,3,1] -> 4x4 outer product / 64."""
    f = torch.as_tensor(1 if taps is None else taps, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(dtype)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:168-208.  x [N,C,H,W]; f [fh,fw] | [taps] | None."""
    n, c, h, w = x.shape
    ux, uy = _pair(up)
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones(1, 1, dtype=torch.float32)
    # zero insertion (:184-186)
    z = x.new_zeros(n, c, h, uy, w, ux)
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(n, c, h * uy, w * ux)
    # pad / crop (:189-190)
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    # filter: conv2d is a correlation, so "convolution" (flip_filter=False) flips (:193-196)
    k = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    if k.ndim == 2:
        z = F.conv2d(z, k[None, None].repeat(c, 1, 1, 1), groups=c)
    else:  # separable: rows then columns (:202-204)
        z = F.conv2d(z, k[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
        z = F.conv2d(z, k[None, None, :, None].repeat(c, 1, 1, 1), groups=c)
    return z[:, :, ::dy, ::dx]


def filter2d(x, f, padding=0, flip_filter=False, gain=1.0):
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1.0):
    ux, uy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = filter_size(f)
    p = [px0 + (fw + ux - 1) // 2, px1 + (fw - ux) // 2, py0 + (fh + uy - 1) // 2, py1 + (fh - uy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1.0):
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = filter_size(f)
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)
