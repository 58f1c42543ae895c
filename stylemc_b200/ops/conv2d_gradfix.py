"""conv2d_gradfix: the reference's ``conv2d`` / ``conv_transpose2d`` entry points (torch_utils/ops/conv2d_gradfix.py:22-43).

In the reference these wrap cuDNN with a hand-written double backward that is only enabled on torch 1.7-1.9
(conv2d_gradfix.py:47-56); on any newer torch they pass straight through to ``F.conv2d`` /
``F.conv_transpose2d``.  Here both names route to the tcgen05 implicit-GEMM kernel through
``conv2d_resample`` (3x3 / 1x1, stride 1 or transposed stride 2 -- the only shapes on the StyleMC path); other
configurations raise, there is no library fallback.
"""
import contextlib

enabled = False                      # conv2d_gradfix.py:22 (kept for drop-in; there is a single CUDA path here)
weight_gradients_disabled = False    # conv2d_gradfix.py:23


@contextlib.contextmanager
def no_weight_gradients():
    """conv2d_gradfix.py:25-31."""
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:                          # an exception inside the block must not leave the switch stuck
        weight_gradients_disabled = old


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """conv2d_gradfix.py:35-38."""
    from . import conv2d_resample
    if _pair(stride) != (1, 1) or _pair(dilation) != (1, 1):
        raise RuntimeError('stylemc_b200 conv2d: only stride 1, dilation 1 is implemented')
    y = conv2d_resample._conv2d(input, weight, padding=_pair(padding), groups=groups, transpose=False)
    return y if bias is None else y + bias.reshape(1, -1, 1, 1)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """conv2d_gradfix.py:40-43."""
    from . import conv2d_resample
    if _pair(stride) != (2, 2) or _pair(dilation) != (1, 1) or _pair(output_padding) != (0, 0) or _pair(padding) != (0, 0):
        raise RuntimeError('stylemc_b200 conv_transpose2d: only stride 2, padding 0 is implemented')
    y = conv2d_resample._conv2d(input, weight, padding=(0, 0), groups=groups, transpose=True)
    return y if bias is None else y + bias.reshape(1, -1, 1, 1)
