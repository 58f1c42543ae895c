"""Host-side launcher for the tcgen05 implicit GEMM (``smc_igemm``): builds the descriptor of
include/stylemc_b200.h from torch tensors.  Operands are fp16 "planes" tensors:

  A : [P, N, H, W, C]   P = 1 (hi only) or 2 (hi, lo) planes of an NHWC activation
  B : [P, rows, C]      P planes of a K-major weight matrix (rows = taps * n_out)

``precision='x1'`` multiplies hi*hi; ``'x3'`` adds hi*lo and lo*hi as extra taps (~21 mantissa bits); ``'x2'`` is the split for an A
operand that has a hi plane only (A_hi*B_hi + A_hi*B_lo): used by the backward GEMMs, whose A operand is a loss-scaled gradient.
"""
import ctypes

import torch

from . import _lib

TAPS_3X3 = [(ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]          # (dy, dx, weight tap)
TAPS_3X3_DGRAD = [(1 - ky, 1 - kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]
TAPS_1X1 = [(0, 0, 0)]


def up2_parity_taps(r, c):
    """Taps of the stride-2 transposed 3x3 conv that land on output parity (r, c): out[2a+r, 2b+c] reads
    in[a - ky//2, b - kx//2] * W[ky, kx] for ky in ({0,2} if r == 0 else {1}), same for kx."""
    kys = (0, 2) if r == 0 else (1,)
    kxs = (0, 2) if c == 0 else (1,)
    return [(-(ky // 2), -(kx // 2), ky * 3 + kx) for ky in kys for kx in kxs]


def up2_dgrad_taps(n_img):
    """dgrad of the same conv reads the gradient parity planes (stacked on the image axis, plane-major):
    g_in[iy, ix] += GP[ky%2][kx%2][iy + ky//2, ix + kx//2] * W[ky, kx]."""
    return [(((ky % 2) * 2 + (kx % 2)) * n_img, ky // 2, kx // 2, ky * 3 + kx) for ky in range(3) for kx in range(3)]


def igemm(A, B, n_img, H, W, n_out, taps, precision='x1', a_plane_stride_imgs=None, b_rows_per_tap=None,
          row_scale=None, post_scale=None, bias=None, noise=None, noise_strides=(0, 0), act=0, alpha=0.2, gain=1.0,
          clamp=-1.0, residual=None, out_f32=None, out_hi=None, out_lo=None, out_raw=None, out_strides=None, out_offset=0,
          tile=None, acc_scale=1.0, acc_chunk_k=0, out_raw_lo=None, rgb_w=None, rgb_acc=None, mask_y=None, mask_y_lo=None, mask_grgb=None,
          problems=None, rgb_part_stride=0):
    """Launch one implicit GEMM.

    A: fp16 tensor viewed as [NA, HA, WA, C] (NA includes the hi/lo planes stacked on the image axis).
    B: fp16 tensor [rowsB, C]; ``taps`` entries are (dn, dy, dx, tap_index) or (dy, dx, tap_index);
       tap_index selects rows tap_index * n_out .. of the hi block, the lo block follows ``b_lo_row`` rows later.
    out_strides: (sn, sh, sw) element strides of the output; default dense NHWC [n_img, H, W, n_out].
    problems: optional [(ntaps, out_offset), ...] (2..4 entries): a problem group (smc_igemm_desc::nprob) -- ``taps`` is the concatenation of
       the problems' tap lists, problem q writes to ``out_offset + problems[q][1]`` (elements).  One launch, the input is read once.
    rgb_part_stride: element stride between the per-N-tile partial-sum images of the fused ToRGB (smc_igemm_epilogue::rgb_snt); 0 = one image.
    mask_y (+ mask_y_lo): fused activation backward (smc_igemm_epilogue::mask_y): out = acc * post_scale * lrelu'(mask_y) * clamp mask,
       optionally + rgb_w . mask_grgb (fp32 NCHW [n_img, 3, H, W]) before the slope.
    """
    assert A.dtype == torch.float16 and B.dtype == torch.float16 and A.is_cuda and B.is_cuda
    assert A.ndim == 4 and B.ndim == 2 and A.is_contiguous() and B.is_contiguous()
    NA, HA, WA, C = A.shape
    d = _lib.IgemmDesc()
    d.A, d.NA, d.HA, d.WA, d.C, d.lda = A.data_ptr(), NA, HA, WA, C, C
    d.B, d.rowsB, d.ldb = B.data_ptr(), B.shape[0], C
    d.n_img, d.H, d.W, d.n_out = n_img, H, W, n_out
    if tile is not None:
        d.tw, d.th, d.tn = tile
    full = []
    for t in taps:
        dn, dy, dx, ti = t if len(t) == 4 else (0,) + tuple(t)
        full.append((dn, dy, dx, ti * n_out))
    if precision == 'x3':
        a_lo = a_plane_stride_imgs if a_plane_stride_imgs is not None else NA // 2
        b_lo = b_rows_per_tap if b_rows_per_tap is not None else B.shape[0] // 2
        full = full + [(dn, dy, dx, br + b_lo) for dn, dy, dx, br in full] + [(dn + a_lo, dy, dx, br) for dn, dy, dx, br in full]
    elif precision == 'x2':
        b_lo = b_rows_per_tap if b_rows_per_tap is not None else B.shape[0] // 2
        full = full + [(dn, dy, dx, br + b_lo) for dn, dy, dx, br in full]
    elif precision != 'x1':
        raise ValueError(precision)
    if len(full) > _lib.MAX_TAPS:
        raise RuntimeError(f'{len(full)} taps > {_lib.MAX_TAPS}')
    d.ntaps = len(full)
    for i, (dn, dy, dx, br) in enumerate(full):
        d.taps[i].dn, d.taps[i].dy, d.taps[i].dx, d.taps[i].brow = dn, dy, dx, br
    e = d.epi
    e.row_scale, e.post_scale, e.bias = _lib.ptr(row_scale), _lib.ptr(post_scale), _lib.ptr(bias)
    e.noise, e.noise_sh, e.noise_sw = _lib.ptr(noise), noise_strides[0], noise_strides[1]
    e.act, e.alpha, e.gain, e.clamp = act, alpha, gain, clamp
    e.residual, e.out_f32 = _lib.ptr(residual), _lib.ptr(out_f32)
    e.out_hi, e.out_lo, e.out_raw = _lib.ptr(out_hi), _lib.ptr(out_lo), _lib.ptr(out_raw)
    if out_strides is None:
        out_strides = (H * W * n_out, W * n_out, n_out)
    e.o_sn, e.o_sh, e.o_sw = out_strides
    e.o_off = out_offset
    e.acc_scale = acc_scale
    e.out_raw_lo, e.rgb_w, e.rgb_acc = _lib.ptr(out_raw_lo), _lib.ptr(rgb_w), _lib.ptr(rgb_acc)
    if rgb_acc is not None:          # NCHW fp32 [n_img, 3, H, W]; rgb_part_stride > 0: one such image per N tile of the kernel, that many elements apart
        assert rgb_acc.dtype == torch.float32 and rgb_acc.is_contiguous() and rgb_w is not None and rgb_w.is_contiguous()
        e.rgb_sn, e.rgb_sj, e.rgb_sh = rgb_acc.stride(0), rgb_acc.stride(1), rgb_acc.stride(2)
        e.rgb_snt = rgb_part_stride
    e.mask_y, e.mask_y_lo, e.mask_grgb = _lib.ptr(mask_y), _lib.ptr(mask_y_lo), _lib.ptr(mask_grgb)
    if mask_grgb is not None:
        assert mask_grgb.dtype == torch.float32 and mask_grgb.is_contiguous() and rgb_w is not None and rgb_w.is_contiguous() and rgb_acc is None
        e.rgb_sn, e.rgb_sj, e.rgb_sh = mask_grgb.stride(0), mask_grgb.stride(1), mask_grgb.stride(2)
    d.acc_chunk_k = acc_chunk_k
    if problems is not None:
        if not 2 <= len(problems) <= 4 or sum(nt for nt, _ in problems) != len(taps):
            raise RuntimeError('problems must be 2..4 (ntaps, out_offset) pairs covering the tap list')
        d.nprob = len(problems)
        for q, (nt, off) in enumerate(problems):
            d.prob_ntaps[q], d.prob_o_off[q] = nt, off
    with torch.cuda.device(A.device):
        if _lib.igemm_hook is not None:
            with _lib.igemm_hook(d, len(full) // {'x3': 3, 'x2': 2}.get(precision, 1)):
                _lib.call('smc_igemm', ctypes.addressof(d), _lib.stream())
        else:
            _lib.call('smc_igemm', ctypes.addressof(d), _lib.stream())


def pow2_prescale(x, target=256.0):
    """Power of two k such that max|x| * k lies in [target/2, target): keeps the lo plane of small weights out of the fp16
    subnormal range.  The GEMM undoes it with ``acc_scale = 1 / k``."""
    import math
    amax = float(x.abs().max())
    if amax == 0.0 or not math.isfinite(amax):
        return 1.0
    return 2.0 ** (math.floor(math.log2(target / amax)))


def prepare_weights(w, two=True, fwd=True, bwd=False, q=False, prescale=False, pad_to=1):
    """Frozen conv / linear weights -> the operand layouts of ``igemm`` in ONE kernel (``smc_prepare_weights``).

    w: [O, I, kh, kw] or [O, I] fp32 CUDA tensor.  Returns (B_fwd [P * T * Op, Ip] or None, B_bwd [P * T * Ip, Op] or None,
    q [O, I] fp32 or None, inv_scale): P = 2 hi/lo planes when ``two``; Op / Ip = channel counts rounded up to ``pad_to``;
    ``prescale`` multiplies the planes by the power of two that brings max|w| into [128, 256) (``pow2_prescale``, found on the device
    by ``smc_grad_scale``) and returns its inverse for ``acc_scale`` -- the one host read of this function."""
    assert w.is_cuda and w.dtype == torch.float32 and w.ndim in (2, 4)
    w = w.contiguous()
    o, i = w.shape[:2]
    t = w.shape[2] * w.shape[3] if w.ndim == 4 else 1
    op, ip = -(-o // pad_to) * pad_to, -(-i // pad_to) * pad_to
    planes = 2 if two else 1
    dev = w.device
    Bf = torch.empty([planes, t * op, ip], dtype=torch.float16, device=dev) if fwd else None
    Bb = torch.empty([planes, t * ip, op], dtype=torch.float16, device=dev) if bwd else None
    qq = torch.empty([o, i], dtype=torch.float32, device=dev) if q else None
    lo = lambda b: _lib.ptr(b[1]) if (b is not None and two) else None
    with torch.cuda.device(dev):
        scale = None
        if prescale:
            amax = torch.zeros(1, dtype=torch.int32, device=dev)
            scale = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.call('smc_grad_scale', _lib.ptr(w), w.numel(), 256.0, _lib.ptr(amax), _lib.ptr(scale), _lib.stream())
        _lib.call('smc_prepare_weights', _lib.ptr(w), o, i, t, op, ip, _lib.ptr(scale), _lib.ptr(Bf), lo(Bf), _lib.ptr(Bb), lo(Bb), _lib.ptr(qq),
                  _lib.stream())
    inv = 1.0 / float(scale) if prescale else 1.0
    return (Bf.reshape(-1, ip) if fwd else None), (Bb.reshape(-1, op) if bwd else None), qq, inv


def split_planes(x, two):
    """fp32 tensor -> stacked fp16 planes [P, ...] (setup-time helper for frozen weights)."""
    hi = x.to(torch.float16)
    if not two:
        return hi.unsqueeze(0).contiguous()
    lo = (x - hi.float()).to(torch.float16)
    return torch.stack([hi, lo]).contiguous()
