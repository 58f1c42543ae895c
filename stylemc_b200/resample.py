"""``unprocess`` of find_direction.py:49-52 as two fused kernels (and their transpose for the backward pass).

The reference does ``clamp(img*127.5+128, 0, 255)`` -> torchvision ``Resize(224, BICUBIC)`` + ``CenterCrop(224)`` -> ``/255`` ->
``(x-mean)/std`` in five ATen passes (find_direction.py:49-52,258; utils.py:90-97).  In this image (torchvision 0.26) Resize on a
tensor is ``F.interpolate(mode='bicubic', antialias=True, align_corners=False)``: a separable filter whose taps depend only on
the output index.  The tap tables are built here on the host exactly like ATen's ``_compute_indices_weights_aa`` (bicubic
a = -0.5, support scaled by the down-sampling factor); the kernels fold denormalise+clamp into the horizontal pass and
/255+normalise into the vertical pass.
"""
import ctypes
import functools

import numpy as np
import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # utils.py:91
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)     # utils.py:92


def _cubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def aa_tables(in_size, out_size):
    """Per output index: first input index, tap count, normalised weights (float32, like ATen computes them for float input).
    Returns (start[out], count[out], weights[out, taps], taps)."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    taps = int(np.ceil(support)) * 2 + 1
    start = np.zeros(out_size, np.int32)
    count = np.zeros(out_size, np.int32)
    wgt = np.zeros((out_size, taps), np.float32)
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([_cubic((j + xmin - center + 0.5) * invscale) for j in range(xsize)], np.float32)
        total = np.float32(w.sum(dtype=np.float32))
        if total != 0:
            w = w / total
        start[i], count[i] = xmin, xsize
        wgt[i, :xsize] = w
    return start, count, wgt, taps


def transpose_tables(start, count, wgt, in_size):
    """Per INPUT index: the outputs that read it and with which weight (for the backward pass)."""
    lists = [[] for _ in range(in_size)]
    for o in range(len(start)):
        for k in range(count[o]):
            lists[start[o] + k].append((o, wgt[o, k]))
    taps = max(1, max(len(l) for l in lists))
    oidx = np.zeros((in_size, taps), np.int32)
    cnt = np.zeros(in_size, np.int32)
    w = np.zeros((in_size, taps), np.float32)
    for i, l in enumerate(lists):
        cnt[i] = len(l)
        for k, (o, v) in enumerate(l):
            oidx[i, k], w[i, k] = o, v
        # the kernels read oidx[i, 0] only: the outputs that read one input are consecutive (monotone windows)
        assert all(o == l[0][0] + k for k, (o, _) in enumerate(l))
    return oidx, cnt, w, taps


def dense_matrix(in_size, out_size):
    """[out, in] resampling matrix (host-side check of the tables against F.interpolate)."""
    start, count, wgt, _ = aa_tables(in_size, out_size)
    m = np.zeros((out_size, in_size), np.float64)
    for o in range(out_size):
        m[o, start[o]:start[o] + count[o]] = wgt[o, :count[o]]
    return m


@functools.lru_cache(maxsize=16)
def _device_tables(in_size, out_size, device):
    start, count, wgt, taps = aa_tables(in_size, out_size)
    oidx, cnt_t, wgt_t, taps_t = transpose_tables(start, count, wgt, in_size)
    dev = torch.device(device)
    t = lambda a: torch.as_tensor(a).to(dev).contiguous()
    return dict(start=t(start), count=t(count), wgt=t(wgt), taps=taps, oidx=t(oidx), count_t=t(cnt_t), wgt_t=t(wgt_t), taps_t=taps_t)


_F3 = ctypes.c_float * 3
_MEAN, _STD = _F3(*CLIP_MEAN), _F3(*CLIP_STD)


MODES = {'unprocess': 1, 'nada': 2}     # find_direction.py:49-52 | clip_loss_nada.py:86-89 ((x + 1) / 2, no clamp, no / 255)


def unprocess_fwd(img, size=224, mode='unprocess'):
    """img [N, 3, R, R] fp32 CUDA ([-1, 1]-ish) -> CLIP-normalised [N, 3, size, size] fp32."""
    _lib.require_cuda(img, 'img')
    n, c, h, w = img.shape
    if c != 3 or h != w:
        raise RuntimeError('unprocess expects square RGB images [N, 3, R, R]')
    img = img.float().contiguous()
    tb = _device_tables(h, size, str(img.device))
    tmp = torch.empty([n * 3, h, size], dtype=torch.float32, device=img.device)
    out = torch.empty([n, 3, size, size], dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.call('smc_resample_fwd', _lib.ptr(img), _lib.ptr(tmp), _lib.ptr(out), _lib.ptr(tb['start']), _lib.ptr(tb['count']),
                  _lib.ptr(tb['wgt']), tb['taps'], n * 3, h, size, MODES[mode], ctypes.addressof(_MEAN), ctypes.addressof(_STD), _lib.stream())
    return out


def unprocess_bwd(g_out, img, unscale=None, mode='unprocess'):
    """Transpose of unprocess_fwd: g_out [N, 3, S, S] -> gradient w.r.t. img [N, 3, R, R].  ``unscale`` is an optional device
    scalar S: the result is divided by it (loss scale carried by g_out)."""
    n, c, size, _ = g_out.shape
    h = img.shape[2]
    tb = _device_tables(h, size, str(img.device))
    g_out = g_out.float().contiguous()
    tmp = torch.empty([n * 3, h, size], dtype=torch.float32, device=img.device)
    gx = torch.empty_like(img)
    with torch.cuda.device(img.device):
        _lib.call('smc_resample_bwd', _lib.ptr(g_out), _lib.ptr(img), _lib.ptr(tmp), _lib.ptr(gx), _lib.ptr(tb['oidx']), _lib.ptr(tb['count_t']),
                  _lib.ptr(tb['wgt_t']), tb['taps_t'], n * 3, h, size, MODES[mode], ctypes.addressof(_STD), _lib.ptr(unscale), _lib.stream())
    return gx


class _Unprocess(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, size, mode='unprocess'):
        ctx.save_for_backward(img)
        ctx.mode = mode
        return unprocess_fwd(img, size, mode)

    @staticmethod
    def backward(ctx, g):
        img, = ctx.saved_tensors
        return unprocess_bwd(g, img.float().contiguous(), mode=ctx.mode), None, None


def unprocess(img, transf=None, mean=None, std=None, size=224):
    """Drop-in for find_direction.unprocess(img, transf, mean, std): the transform, mean and std arguments are accepted for
    signature compatibility; the kernel implements Resize(224, BICUBIC)+CenterCrop(224) on square inputs with CLIP's constants."""
    return _Unprocess.apply(img, size)


def nada_preprocess(img, size=224):
    """``CLIPLoss.preprocess`` of clip_loss_nada.py:86-89 on a square GAN output: Normalize(mean -1, std 2) (no clamp), Resize(224, BICUBIC) +
    CenterCrop(224), Normalize(CLIP mean, std).  Differentiable."""
    return _Unprocess.apply(img, size, 'nada')
