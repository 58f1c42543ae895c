"""Host-side mirror of the network modules the reference reaches through its pickles.

``training/networks.py`` of NVlabs/stylegan2-ada-pytorch is not in the reference tree (it is exec()'d out of the network
pickle, torch_utils/persistence.py:179-227); the reference touches these classes through the attribute names fixed by
legacy.py:173-202 and the calls ``layer(x, w, noise_mode=, fused_modconv=, gain=)`` / ``torgb(x, w, fused_modconv=)`` at
utils.py:32-47.  This module provides the same classes, parameter names and call conventions for the S-space path
(FullyConnectedLayer, SynthesisLayer, ToRGBLayer, SynthesisBlock, SynthesisNetwork, Generator) and the op
``modulated_conv2d``; their arithmetic runs on the package's CUDA kernels (ops.*, smc_igemm).  Layer-by-layer calls go
through the op-level NCHW API; the fast path is ``utils.generate_image`` -> ``synthesis.SynthesisEngine``, which executes a
whole network as fused NHWC kernels.
"""
import math

import numpy as np
import torch

from . import _lib, gemm
from .ops import bias_act, conv2d_resample, fma, upfirdn2d


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """[UPSTREAM modulated_conv2d; SURVEY.md section 8 A1]  x [N,I,H,W], weight [O,I,kh,kw] (frozen), styles [N,I].

    Both values of ``fused_modconv`` produce the same result; the computation is always "scale activations, shared-weight
    conv, scale by the demodulation coefficients" (no per-sample weight tensor is materialised).  Differentiable w.r.t.
    ``x`` and ``styles``."""
    n = x.shape[0]
    cout, cin, kh, kw = weight.shape
    assert x.shape[1] == cin and tuple(styles.shape) == (n, cin)
    w32 = weight.float()
    if x.dtype == torch.float16 and demodulate:    # upstream pre-normalisation of the fp16 branch (cancels in exact arithmetic)
        w32 = w32 * (1 / np.sqrt(cin * kh * kw) / w32.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    dcoefs = None
    if demodulate:
        q = w32.square().sum(dim=[2, 3])                                                   # [O, I]
        dcoefs = (styles.float().square() @ q.t() + 1e-8).rsqrt()                          # [N, O]
    xm = x * styles.to(x.dtype).reshape(n, cin, 1, 1)
    with conv2d_resample.conv2d_gradfix.no_weight_gradients():
        y = conv2d_resample.conv2d_resample(xm, w32.to(x.dtype).detach(), f=resample_filter, up=up, down=down, padding=padding,
                                            flip_weight=flip_weight)
    if demodulate and noise is not None:
        return fma.fma(y, dcoefs.to(y.dtype).reshape(n, cout, 1, 1), noise.to(y.dtype))
    if demodulate:
        return y * dcoefs.to(y.dtype).reshape(n, cout, 1, 1)
    if noise is not None:
        return y + noise.to(y.dtype)
    return y


class FullyConnectedLayer(torch.nn.Module):
    """[UPSTREAM] y = x @ (W * lr/sqrt(in)).T + b * lr.  Only used to turn W into S once per seed (utils.py:123-158)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1.0, bias_init=0.0):
        super().__init__()
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn(out_features, in_features) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], float(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / math.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype) * self.bias_gain if self.bias_gain != 1 else b.to(x.dtype)
        y = x.matmul(w.t())
        if self.activation == 'linear':
            return y if b is None else y + b
        return bias_act.bias_act(y, b, act=self.activation)


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                 resample_filter=(1, 3, 3, 1), conv_clamp=None, channels_last=False):
        super().__init__()
        self.resolution, self.up, self.use_noise = resolution, up, use_noise
        self.activation, self.conv_clamp = activation, conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(list(resample_filter)))
        self.padding = kernel_size // 2
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        if use_noise:
            self.register_buffer('noise_const', torch.randn(resolution, resolution))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1):
        assert noise_mode in ('random', 'const', 'none')
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        x = modulated_conv2d(x, self.weight, styles, noise=noise, up=self.up, padding=self.padding,
                             resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv)
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return bias_act.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=clamp)


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.weight_gain = 1 / math.sqrt(in_channels * kernel_size ** 2)

    def forward(self, x, w, fused_modconv=True):
        styles = self.affine(w) * self.weight_gain
        x = modulated_conv2d(x, self.weight, styles, demodulate=False, fused_modconv=fused_modconv)
        return bias_act.bias_act(x, self.bias.to(x.dtype), clamp=self.conv_clamp)


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=(1, 3, 3, 1), conv_clamp=None, use_fp16=False, fp16_channels_last=False, **layer_kwargs):
        super().__init__()
        if architecture != 'skip':
            raise RuntimeError("only the 'skip' architecture of the FFHQ config-f networks is implemented")
        self.in_channels, self.w_dim, self.resolution = in_channels, w_dim, resolution
        self.img_channels, self.is_last, self.architecture = img_channels, is_last, architecture
        self.use_fp16 = use_fp16
        self.channels_last = use_fp16 and fp16_channels_last
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(list(resample_filter)))
        self.num_conv = self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn(out_channels, resolution, resolution))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim, resolution, up=2, resample_filter=resample_filter,
                                        conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim, resolution, conv_clamp=conv_clamp, **layer_kwargs)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim, conv_clamp=conv_clamp)
        self.num_torgb += 1

    def forward(self, x, img, ws, force_fp32=True, fused_modconv=None, **layer_kwargs):
        """[UPSTREAM SynthesisBlock.forward] driven from W; ``utils.block_forward`` (utils.py:13-53) is the S-driven twin."""
        w_iter = iter(ws.unbind(dim=1))
        if self.in_channels == 0:
            x = self.const.unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        else:
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        if img is not None:
            img = upfirdn2d.upsample2d(img, self.resample_filter)
        y = self.torgb(x, next(w_iter), fused_modconv=fused_modconv).to(torch.float32)
        img = img.add_(y) if img is not None else y
        return x, img


class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim=512, img_resolution=1024, img_channels=3, channel_base=32768, channel_max=512, num_fp16_res=0,
                 **block_kwargs):
        super().__init__()
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        log2 = int(np.log2(img_resolution))
        self.img_resolution_log2 = log2
        self.block_resolutions = [2 ** i for i in range(2, log2 + 1)]
        channels = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        fp16_resolution = max(2 ** (log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            block = SynthesisBlock(channels[res // 2] if res > 4 else 0, channels[res], w_dim=w_dim, resolution=res,
                                   img_channels=img_channels, is_last=(res == img_resolution), use_fp16=(res >= fp16_resolution),
                                   **block_kwargs)
            self.num_ws += block.num_conv
            if res == img_resolution:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward(self, ws, **block_kwargs):
        """[UPSTREAM SynthesisNetwork.forward]: W -> image through the module-level path (generate_fromS.py:98 ``G.synthesis(w)``)."""
        block_ws, w_idx = [], 0
        ws = ws.to(torch.float32)
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            block_ws.append(ws.narrow(1, w_idx, block.num_conv + block.num_torgb))
            w_idx += block.num_conv
        x = img = None
        for res, cur in zip(self.block_resolutions, block_ws):
            x, img = getattr(self, f'b{res}')(x, img, cur, **block_kwargs)
        return img


class MappingNetwork(torch.nn.Module):
    """[UPSTREAM MappingNetwork, c_dim = 0] z -> W+ (generate_w.py:48-50, once per seed set): second-moment normalisation, 8
    FullyConnectedLayers (lrelu through the ``bias_act`` kernel, lr multiplier 0.01; kwargs of legacy.py:129-136), broadcast to
    ``num_ws`` rows, truncation ``w_avg.lerp(w, psi)``.  Attribute names ``fc{i}`` / ``w_avg`` as legacy.py:175-181."""

    def __init__(self, z_dim=512, w_dim=512, num_ws=18, num_layers=8, lr_multiplier=0.01):
        super().__init__()
        self.z_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, w_dim, num_ws, num_layers
        for i in range(num_layers):
            setattr(self, f'fc{i}', FullyConnectedLayer(z_dim if i == 0 else w_dim, w_dim, activation='lrelu', lr_multiplier=lr_multiplier))
        self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1, truncation_cutoff=None):
        x = z.to(torch.float32)
        x = x * (x.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
        for i in range(self.num_layers):
            x = getattr(self, f'fc{i}')(x)
        x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            if truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class Generator(torch.nn.Module):
    """Stand-in for the unpickled ``G_ema``: ``.synthesis`` is the accelerated path; ``.mapping`` (z -> W+, once per seed set,
    generate_w.py:48-51) is built after it, as upstream, and only when asked for."""

    def __init__(self, mapping=False, **synthesis_kwargs):
        super().__init__()
        self.synthesis = SynthesisNetwork(**synthesis_kwargs)
        self.z_dim, self.c_dim, self.w_dim = 512, 0, self.synthesis.w_dim
        if mapping:
            self.mapping = MappingNetwork(z_dim=self.z_dim, w_dim=self.w_dim, num_ws=self.synthesis.num_ws)


def make_generator(img_resolution, seed=0, channel_base=32768, channel_max=512, conv_clamp=256, noise_strength=0.1,
                   torgb_scale=0.25, mapping=False):
    """Random-init FFHQ config-f style generator (BASELINE.json: random-init weights; upstream init N(0,1) weights, zero biases,
    affine bias 1) with the noise path switched on (noise_strength) and ToRGB weights scaled so images mostly lie in [-1, 1]."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        G = Generator(mapping=mapping, w_dim=512, img_resolution=img_resolution, img_channels=3, channel_base=channel_base,
                      channel_max=channel_max, num_fp16_res=0, conv_clamp=conv_clamp)
        if mapping:
            with torch.no_grad():
                G.mapping.w_avg.copy_(0.1 * torch.randn(512))       # a trained network carries the running mean of W here
    finally:
        torch.random.set_rng_state(state)
    with torch.no_grad():
        for res in G.synthesis.block_resolutions:
            blk = getattr(G.synthesis, f'b{res}')
            blk.torgb.weight.mul_(torgb_scale)
            for name in ('conv0', 'conv1'):
                if hasattr(blk, name):
                    getattr(blk, name).noise_strength.fill_(noise_strength)
    return G.eval().requires_grad_(False)
