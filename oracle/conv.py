"""Oracle: resampling convolution and the modulated convolution  (TEST INFRASTRUCTURE ONLY).

* ``conv2d_resample``  restates ``torch_utils/ops/conv2d_resample.py:59-154`` (dispatch) and
  ``_conv2d_wrapper`` (:29-54); the convs themselves are the pass-through of
  ``conv2d_gradfix.py:35-43`` on torch 2.x, i.e. plain F.conv2d / F.conv_transpose2d.
* ``fma`` restates ``torch_utils/ops/fma.py:15-58`` (a*b+c).
* ``modulated_conv2d`` restates NVlabs stylegan2-ada-pytorch ``training/networks.py``
  (absent from the reference tree -- PARITY UNPINNED; semantics per SURVEY.md section 8 A1 and the
  call sites ``utils.py:32-47``): w' = W*s, d = rsqrt(sum w'^2 + 1e-8), fused grouped conv or
  scale-activations form.
"""
import torch
import torch.nn.functional as F

from . import fir


def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    # conv2d_resample.py:29-54: F.conv2d correlates; flip_weight=False asks for true convolution.
    if not flip_weight:
        w = w.flip([2, 3])
    if transpose:
        return F.conv_transpose2d(x, w, stride=stride, padding=padding, groups=groups)
    return F.conv2d(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:59-154."""
    cout, cin_g, kh, kw = w.shape
    fw, fh = fir.filter_size(f)
    px0, px1, py0, py1 = fir._pad4(padding)
    if up > 1:  # :84-88
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
        x = fir.upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:  # :102-105
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return fir.upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:  # :108-111
        x = fir.upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:  # :114-133 transposed strided conv then FIR
        if groups == 1:
            wt = w.transpose(0, 1)
        else:
            wt = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2)
            wt = wt.reshape(groups * cin_g, cout // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv(x, wt, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=not flip_weight)
        x = fir.upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2,
                          flip_filter=flip_filter)
        if down > 1:
            x = fir.upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:  # :136-138
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    # generic fallback :141-145
    x = fir.upfirdn2d(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2,
                      flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = fir.upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


def fma(a, b, c):
    """fma.py:15-23 (forward value; autograd supplies the broadcast-aware backward of :25-58)."""
    return torch.addcmul(c, a, b)


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused_modconv=True):
    """[UPSTREAM training/networks.py modulated_conv2d]  x [N,I,H,W], weight [O,I,kh,kw], styles [N,I]."""
    n = x.shape[0]
    cout, cin, kh, kw = weight.shape
    if x.dtype == torch.float16 and demodulate:  # fp16 pre-normalisation branch
        weight = weight * (1 / (cin * kh * kw) ** 0.5 / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    wmod = dcoef = None
    if demodulate or fused_modconv:
        wmod = weight.unsqueeze(0) * styles.reshape(n, 1, cin, 1, 1)            # [N,O,I,kh,kw]
    if demodulate:
        dcoef = (wmod.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()               # [N,O]
    if demodulate and fused_modconv:
        wmod = wmod * dcoef.reshape(n, cout, 1, 1, 1)

    if not fused_modconv:  # scale activations before and after a shared-weight conv
        x = x * styles.to(x.dtype).reshape(n, cin, 1, 1)
        x = conv2d_resample(x, weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                            flip_weight=flip_weight)
        if demodulate and noise is not None:
            return fma(x, dcoef.to(x.dtype).reshape(n, cout, 1, 1), noise.to(x.dtype))
        if demodulate:
            return x * dcoef.to(x.dtype).reshape(n, cout, 1, 1)
        if noise is not None:
            return x + noise.to(x.dtype)
        return x

    # one grouped conv with groups = batch
    x = x.reshape(1, n * cin, *x.shape[2:])
    x = conv2d_resample(x, wmod.reshape(n * cout, cin, kh, kw).to(x.dtype), f=resample_filter, up=up, down=down,
                        padding=padding, groups=n, flip_weight=flip_weight)
    x = x.reshape(n, cout, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x
