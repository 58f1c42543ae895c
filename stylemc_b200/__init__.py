"""stylemc_b200 -- B200-native (sm_100a) implementation of StyleMC's S-space synthesis + CLIP-loss hot path.

Host side mirrors the reference's operator interface (``torch_utils.ops``, ``utils.generate_image``,
``clip_loss.CLIPLoss``, the find_direction step); compute runs in hand-written CUDA kernels reached through
the C ABI of ``libstylemc_b200.so`` (include/stylemc_b200.h).  There is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401

__version__ = '0.1.0'
