"""Drop-in mirror of the reference's ``torch_utils.ops`` package for the hot path:
bias_act, upfirdn2d, conv2d_resample, conv2d_gradfix, fma (reference torch_utils/ops/*.py)."""
from . import bias_act, conv2d_gradfix, conv2d_resample, fma, upfirdn2d  # noqa: F401
from . import custom_ops  # noqa: F401  (registers torch.ops.stylemc_b200.bias_act / .upfirdn2d)
