"""Oracle: y = clamp(act(x + b) * gain)  (TEST INFRASTRUCTURE ONLY).

Restates ``torch_utils/ops/bias_act.py:93-123`` (``_bias_act_ref``) and the activation table
``bias_act.py:23-33`` (name -> func, default alpha, default gain, plugin index, which tensor the
gradient is expressed in).
"""
import math
from collections import namedtuple

import torch
import torch.nn.functional as F

Spec = namedtuple('Spec', 'fn def_alpha def_gain cuda_idx ref has_2nd_grad')

ACTIVATIONS = {
    'linear':   Spec(lambda x, a: x,                      0.0, 1.0,          1, '',  False),
    'relu':     Spec(lambda x, a: F.relu(x),              0.0, math.sqrt(2), 2, 'y', False),
    'lrelu':    Spec(lambda x, a: F.leaky_relu(x, a),     0.2, math.sqrt(2), 3, 'y', False),
    'tanh':     Spec(lambda x, a: torch.tanh(x),          0.0, 1.0,          4, 'y', True),
    'sigmoid':  Spec(lambda x, a: torch.sigmoid(x),       0.0, 1.0,          5, 'y', True),
    'elu':      Spec(lambda x, a: F.elu(x),               0.0, 1.0,          6, 'y', True),
    'selu':     Spec(lambda x, a: F.selu(x),              0.0, 1.0,          7, 'y', True),
    'softplus': Spec(lambda x, a: F.softplus(x),          0.0, 1.0,          8, 'y', True),
    'swish':    Spec(lambda x, a: torch.sigmoid(x) * x,   0.0, math.sqrt(2), 9, 'x', True),
}


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """bias_act.py:93-123."""
    spec = ACTIVATIONS[act]
    alpha = float(spec.def_alpha if alpha is None else alpha)
    gain = float(spec.def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if b is not None:
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + b.reshape(shape)
    x = spec.fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x
