"""The plugin boundary as ``torch.library`` custom operators.

The reference reaches its two native plugins through pybind modules built by ``torch_utils/custom_ops.py:46-124``:
``_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp) -> Tensor`` (bias_act.cpp:32; callers bias_act.py:153,182,201)
and ``_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain) -> Tensor`` (upfirdn2d.cpp:16; caller
upfirdn2d.py:237-240).  Here the same two calls are ``torch.ops.stylemc_b200.bias_act`` / ``torch.ops.stylemc_b200.upfirdn2d``: same
argument lists (an absent tensor is ``None`` or, as in the reference, an empty tensor: ``_null_tensor``, bias_act.py:39), CUDA implementation =
one launch through the C ABI (``smc_bias_act`` / ``smc_upfirdn2d``), a fake (meta) implementation for tracing, and no CPU kernel -- a
CPU tensor raises.  Like the pybind functions they carry no autograd formula: differentiation lives one level up, in the
``autograd.Function`` pairs of ``ops/bias_act.py`` and ``ops/upfirdn2d.py`` (as in the reference, bias_act.py:129-210, upfirdn2d.py:214-268),
which call the same launchers.
"""
from typing import Optional

import torch

from . import bias_act as _bias_act
from . import upfirdn2d as _upfirdn2d

_SPEC_BY_IDX = {spec.cuda_idx: spec for spec in _bias_act.activation_funcs.values()}


@torch.library.custom_op('stylemc_b200::bias_act', mutates_args=(), device_types='cuda')
def bias_act(x: torch.Tensor, b: Optional[torch.Tensor], xref: Optional[torch.Tensor], yref: Optional[torch.Tensor],
             dy: Optional[torch.Tensor], grad: int, dim: int, act: int, alpha: float, gain: float, clamp: float) -> torch.Tensor:
    if act not in _SPEC_BY_IDX:
        raise RuntimeError(f'bias_act: unknown activation index {act}')
    if grad not in (0, 1, 2):
        raise RuntimeError('grad must be 0, 1 or 2')                               # bias_act.cpp:44
    present = lambda t: t if (t is not None and t.numel() > 0) else None        # the reference passes an empty tensor for "absent" (bias_act.py:39)
    return _bias_act._launch(x, present(b), present(xref), present(yref), present(dy), grad, dim, _SPEC_BY_IDX[act], float(alpha), float(gain),
                             float(clamp))


@bias_act.register_fake
def _(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
    return torch.empty_like(x)


@torch.library.custom_op('stylemc_b200::upfirdn2d', mutates_args=(), device_types='cuda')
def upfirdn2d(x: torch.Tensor, f: torch.Tensor, upx: int, upy: int, downx: int, downy: int, padx0: int, padx1: int, pady0: int, pady1: int,
              flip: bool, gain: float) -> torch.Tensor:
    if min(upx, upy, downx, downy) < 1:
        raise RuntimeError('up and down factors must be >= 1')                     # upfirdn2d.cpp:27-28
    return _upfirdn2d._launch(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)


@upfirdn2d.register_fake
def _(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    n, c, ih, iw = x.shape
    ow = (iw * upx + padx0 + padx1 - f.shape[1] + downx) // downx                  # upfirdn2d.cpp:32-33
    oh = (ih * upy + pady0 + pady1 - f.shape[0] + downy) // downy
    fmt = torch.channels_last if (x.stride(1) == 1 and c > 1) else torch.contiguous_format
    return torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device, memory_format=fmt)
