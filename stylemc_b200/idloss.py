"""The identity term of the find_direction objective (find_direction.py:179-180): ``id_loss.IDLoss`` on an IR-SE50 ArcFace backbone.

Reference: ``id_loss/id_loss.py:8-39`` (pool to 256 -> crop [35:223, 32:220] -> pool to 112 -> backbone -> 1 - <f(y_hat), f(y)>, mean over the
batch, f(y) detached), ``id_loss/model_irse.py:10-49`` (Backbone(112, 50, 'ir_se')), ``id_loss/helpers.py:17-119`` (units, SEModule, l2_norm).
Parameters arrive as a flat dict under the reference's ``state_dict`` keys (what ``torch.load('id_loss/model_ir_se50.pth')`` returns).

This is the OP-LEVEL formulation of SURVEY.md section 8 (f4): every layer is a call into a repo kernel with a hand-written backward --
convolutions and the final Linear on the tcgen05 implicit GEMM (``smc_igemm``, split-fp16 operands with promoted accumulation, frozen
weights prepared once by ``smc_prepare_weights``), BatchNorm (eval mode: an affine map) / SE gate / residual add on ``smc_fma`` and
``smc_fma_reduce``, PReLU on ``smc_prelu``, ReLU / sigmoid on ``smc_bias_act``, the two adaptive pools on ``smc_adaptive_avg_pool``, and the
normalised dot product on ``smc_clip_loss`` with one target vector per sample.  It is not fused like the synthesis engine (activations
travel as NCHW fp32 between layers); G and the backbone are frozen, so only input gradients exist.
"""
import torch

from . import _lib, gemm
from .ops import bias_act, fma
from .ops.conv2d_resample import _pack, _unpack

UNITS = [(cin, depth, 2 if j == 0 else 1) for cin0, depth, n in ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3))
         for j, cin in enumerate([cin0] + [depth] * (n - 1))]                         # helpers.py:28-38 (num_layers 50)
TAPS = {1: [(0, 0, 0)], 3: gemm.TAPS_3X3}
TAPS_DGRAD = {1: [(0, 0, 0)], 3: gemm.TAPS_3X3_DGRAD}


class _FrozenConv:
    """Conv2d(cin, cout, k, stride, padding=k // 2, bias=False) with frozen weights: operand planes prepared once."""

    def __init__(self, w, stride=1):
        w = w.detach().float().contiguous()
        if w.ndim == 2:
            w = w[:, :, None, None]
        self.cout, self.cin, self.k = w.shape[0], w.shape[1], w.shape[2]
        self.stride = stride
        self.cin_p, self.cout_p = -(-self.cin // 32) * 32, -(-self.cout // 32) * 32
        # the planes carry a power-of-two prescale (max |w| -> [128, 256), undone by acc_scale): He-scaled weights of 0.006 .. 0.06 would put
        # the lo plane into the fp16 subnormals (absolute error 3e-8 = 5e-6 of a weight of the final Linear)
        self.B_fwd, self.B_bwd, _, self.inv = gemm.prepare_weights(w, two=True, fwd=True, bwd=True, pad_to=32, prescale=True)

    def __call__(self, x):
        return _ConvFn.apply(x, self)


def _igemm_nchw(x, B, n_out_p, n_out, taps, acc_chunk_k, acc_scale):
    """NCHW fp32 -> NCHW fp32 through one split-precision implicit GEMM (A planes packed NHWC, output unpacked)."""
    n, c, h, w = x.shape
    A = _pack(x, 2)
    y = torch.empty([n, h, w, n_out_p], dtype=torch.float32, device=x.device)
    gemm.igemm(A, B, n, h, w, n_out_p, taps, precision='x3', acc_chunk_k=acc_chunk_k, out_f32=y, acc_scale=acc_scale)
    return _unpack(y, n_out, torch.float32)


# K elements per promoted accumulation chain.  Forward: 64 (chains of 4 MMAs: ~1e-7 forward error) -- the network has 48 PReLU / ReLU kinks
# in series, and a unit whose pre-activation lies within the forward error of 0 takes the other slope in the backward pass (DESIGN.md
# section 5); with 512 the image gradient of the golden sits at 9.9e-4.  Backward: 512.
ACC_K_FWD, ACC_K_BWD = 64, 512


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, L):
        ctx.L, ctx.in_hw = L, x.shape[2:]
        x = x.float()
        if L.stride == 2 and L.k == 1:
            x = x[:, :, ::2, ::2]                                   # a strided 1x1 conv reads the even pixels only
        y = _igemm_nchw(x.contiguous(), L.B_fwd, L.cout_p, L.cout, TAPS[L.k], ACC_K_FWD, L.inv)
        if L.stride == 2 and L.k == 3:
            y = y[:, :, ::2, ::2].contiguous()                      # padding 1, stride 2: the even positions of the stride-1 result
        return y

    @staticmethod
    def backward(ctx, gy):
        L = ctx.L
        h, w = ctx.in_hw
        gy = gy.float()
        if L.stride == 2 and L.k == 3:                              # transpose of the sub-sampling: zeros at the odd positions
            full = torch.zeros([gy.shape[0], gy.shape[1], h, w], dtype=torch.float32, device=gy.device)
            full[:, :, ::2, ::2] = gy
            gy = full
        gx = _igemm_nchw(gy.contiguous(), L.B_bwd, L.cin_p, L.cin, TAPS_DGRAD[L.k], ACC_K_BWD, L.inv)
        if L.stride == 2 and L.k == 1:
            full = torch.zeros([gx.shape[0], gx.shape[1], h, w], dtype=torch.float32, device=gx.device)
            full[:, :, ::2, ::2] = gx
            gx = full
        return gx, None


class _PreluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        x = x.contiguous()
        ctx.save_for_backward(x, alpha)
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call('smc_prelu', _lib.ptr(x), None, _lib.ptr(alpha), _lib.ptr(y), x.numel(), x.shape[2] * x.shape[3], x.shape[1], _lib.stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        x, alpha = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call('smc_prelu', _lib.ptr(x), _lib.ptr(dy), _lib.ptr(alpha), _lib.ptr(dx), x.numel(), x.shape[2] * x.shape[3], x.shape[1], _lib.stream())
        return dx, None


class _PoolFn(torch.autograd.Function):
    """AdaptiveAvgPool2d(out) of the window [y0:y1, x0:x1] (id_loss.py:12-13,19-22)."""

    @staticmethod
    def forward(ctx, x, window, out):
        x = x.float().contiguous()
        n, c, h, w = x.shape
        y0, y1, x0, x1 = window
        ctx.geom = (n * c, h, w, y0, x0, y1 - y0, x1 - x0, out, out)
        y = torch.empty([n, c, out, out], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call('smc_adaptive_avg_pool', _lib.ptr(x), _lib.ptr(y), *ctx.geom, 0, _lib.stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        planes, h, w = ctx.geom[:3]
        dy = dy.float().contiguous()
        dx = torch.empty([dy.shape[0], dy.shape[1], h, w], dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            _lib.call('smc_adaptive_avg_pool', _lib.ptr(dy), _lib.ptr(dx), *ctx.geom, 1, _lib.stream())
        return dx, None, None


def face_crop(x):
    """id_loss.py:18-22."""
    if x.shape[2] != 256:
        x = _PoolFn.apply(x, (0, x.shape[2], 0, x.shape[3]), 256)
    return _PoolFn.apply(x, (35, 223, 32, 220), 112)


class IRSE50:
    """``Backbone(input_size=112, num_layers=50, mode='ir_se')`` in eval mode (id_loss.py:11,14), frozen, on ``device``.
    ``__call__`` returns the 512 features BEFORE the final l2_norm (model_irse.py:49): the normalisation lives in the loss kernel."""

    def __init__(self, params, device='cuda'):
        self.device = torch.device(device)
        f = lambda k: params[k].detach().to(self.device, torch.float32).contiguous()

        def bn(key, extra_bias=None):          # eval-mode BatchNorm = per-channel affine map (scale, shift), reshaped for broadcasting
            scale = f(key + '.weight') / (f(key + '.running_var') + 1e-5).sqrt()
            shift = f(key + '.bias') - f(key + '.running_mean') * scale
            if extra_bias is not None:
                shift = shift + extra_bias * scale
            return scale.view(1, -1, 1, 1).contiguous(), shift.view(1, -1, 1, 1).contiguous()

        self.conv_in = _FrozenConv(f('input_layer.0.weight'))
        self.bn_in, self.prelu_in = bn('input_layer.1'), f('input_layer.2.weight')
        self.units = []
        for u, (cin, depth, stride) in enumerate(UNITS):
            b = f'body.{u}.'
            self.units.append(dict(
                stride=stride,
                shortcut=None if cin == depth else (_FrozenConv(f(b + 'shortcut_layer.0.weight'), stride), bn(b + 'shortcut_layer.1')),
                bn0=bn(b + 'res_layer.0'), conv1=_FrozenConv(f(b + 'res_layer.1.weight')), prelu=f(b + 'res_layer.2.weight'),
                conv2=_FrozenConv(f(b + 'res_layer.3.weight'), stride), bn1=bn(b + 'res_layer.4'),
                fc1=_FrozenConv(f(b + 'res_layer.5.fc1.weight')), fc2=_FrozenConv(f(b + 'res_layer.5.fc2.weight'))))
        self.bn_out = bn('output_layer.0')
        self.linear = _FrozenConv(f('output_layer.3.weight'))                                   # Linear(512 * 7 * 7, 512) as a 1x1 conv
        self.bn_feat = bn('output_layer.4', extra_bias=f('output_layer.3.bias'))                # (W x + b) * scale + shift
        self.zero = torch.zeros([], dtype=torch.float32, device=self.device)

    def __call__(self, x):
        """x [N, 3, 112, 112] fp32 CUDA -> [N, 512]."""
        x = _PreluFn.apply(fma.fma(self.conv_in(x), *self.bn_in), self.prelu_in)
        for U in self.units:
            if U['shortcut'] is None:
                shortcut = x[:, :, ::U['stride'], ::U['stride']] if U['stride'] > 1 else x        # MaxPool2d(1, stride), helpers.py:98
            else:
                shortcut = fma.fma(U['shortcut'][0](x), *U['shortcut'][1])
            r = fma.fma(x, *U['bn0'])
            r = _PreluFn.apply(U['conv1'](r), U['prelu'])
            r = fma.fma(U['conv2'](r), *U['bn1'])
            inv_hw = torch.full([], 1.0 / (r.shape[2] * r.shape[3]), dtype=torch.float32, device=self.device)
            s = fma.fma(fma._FmaReduce.apply(r, None, torch.Size([r.shape[0], r.shape[1], 1, 1])), inv_hw, self.zero)   # SEModule: global mean
            s = bias_act.bias_act(U['fc1'](s), act='relu', gain=1)                      # torch.nn.ReLU: bias_act's relu defaults to gain sqrt(2)
            s = bias_act.bias_act(U['fc2'](s), act='sigmoid')
            x = fma.fma(r, s, shortcut)                                                           # gate and residual add in one kernel (strided shortcut read)
        x = fma.fma(x, *self.bn_out)
        x = self.linear(x.reshape(x.shape[0], -1, 1, 1))                                          # Flatten: index c * 49 + h * 7 + w
        return fma.fma(x, *self.bn_feat).reshape(x.shape[0], -1)


class IDLoss:
    """``id_loss.IDLoss`` (id_loss.py:8-39) with the parameters injected.  ``__call__(y_hat, y)`` returns ``(loss, 0)`` like the reference
    (differentiable w.r.t. ``y_hat``); ``loss_and_grad`` is what ``DirectionFinder`` uses."""

    def __init__(self, params, device='cuda'):
        self.facenet = IRSE50(params, device)
        self.device = self.facenet.device

    def extract_feats(self, x, normalize=True):
        """id_loss.py:18-24; ``normalize=False`` returns the features before l2_norm."""
        f = self.facenet(face_crop(x.to(self.device, torch.float32)))
        return f / f.norm(dim=1, keepdim=True) if normalize else f

    def loss_and_grad(self, y_hat, y, coef=1.0, inv_count=None, gscale_target=64.0):
        """-> (loss_part, d loss / d y_hat): loss = coef * inv_count * sum_n (1 - cos(f(y_hat_n), f(y_n))) = coef * n * inv_count + loss_part,
        f(y) detached (id_loss.py:30); inv_count defaults to 1 / batch (the mean of id_loss.py:39)."""
        n = y_hat.shape[0]
        inv_count = 1.0 / n if inv_count is None else inv_count
        with torch.no_grad():
            fy = self.extract_feats(y, normalize=False).contiguous()
        with torch.enable_grad():                    # (also when called from inside an autograd.Function.forward)
            img = y_hat.detach().to(self.device, torch.float32).requires_grad_(True)
            fh = self.extract_feats(img, normalize=False).contiguous()
        part = torch.empty(1, dtype=torch.float32, device=self.device)
        d_f = torch.empty_like(fh)
        gscale = torch.ones(1, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call('smc_clip_loss', _lib.ptr(torch.zeros_like(fh)), _lib.ptr(fh), _lib.ptr(fy), _lib.ptr(part), _lib.ptr(d_f), n, fh.shape[1],
                      float(coef), float(inv_count), _lib.ptr(gscale), float(gscale_target), 0, fh.shape[1], _lib.stream())
        g, = torch.autograd.grad(fh, img, d_f)           # d_f carries the power-of-two loss scale: fp16 operand planes of the dgrad GEMMs
        return part, g / gscale

    def __call__(self, y_hat, y):
        return _IdLossFn.apply(y_hat, y, self), 0.0


class _IdLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, y, mod):
        part, g = mod.loss_and_grad(y_hat, y)
        ctx.save_for_backward(g)
        return (1.0 + part).reshape([])

    @staticmethod
    def backward(ctx, d):
        g, = ctx.saved_tensors
        return g * d, None, None
