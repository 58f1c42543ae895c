"""Oracle: StyleGAN2 synthesis network driven from S-space  (TEST INFRASTRUCTURE ONLY).

Layer classes restate NVlabs stylegan2-ada-pytorch ``training/networks.py`` (FullyConnectedLayer,
SynthesisLayer, ToRGBLayer, SynthesisBlock, SynthesisNetwork) -- that file is NOT in the reference
tree (it is exec()'d out of network pickles, ``torch_utils/persistence.py:179-227``), so these are
PARITY UNPINNED against upstream; attribute names/shapes follow ``legacy.py:173-202`` and the
attribute reads of ``utils.py:13-53,100-157`` so that the reference's own ``utils.generate_image``
can drive these modules (that is how ``pin_reference.py`` pins the driver functions below).

Driver functions restate the reference's in-tree S-space driver:
``split_ws`` utils.py:77-87, ``get_temp_shapes`` :100-120, ``get_styles`` :123-158,
``block_forward`` :13-53, ``generate_image`` :161-216 (blending branches :189-205 are out of scope).
"""
import math

import torch

from . import act, conv, fir

N_STYLE_ROWS = 26      # find_direction.py:38
STYLE_WIDTH = 512      # utils.py:125


class FullyConnectedLayer(torch.nn.Module):
    """[UPSTREAM] y = x @ (W*gain).T + b*lr ; W ~ N(0,1)/lr, gain = lr/sqrt(in)."""

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
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return act.bias_act(x.matmul(w.t()), b, act=self.activation)


class SynthesisLayer(torch.nn.Module):
    """[UPSTREAM] modulated 3x3 conv (+up=2 FIR) + noise + bias + lrelu*sqrt2 (+clamp)."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True,
                 activation='lrelu', resample_filter=(1, 3, 3, 1), conv_clamp=None, channels_last=False):
        super().__init__()
        self.resolution, self.up, self.use_noise = resolution, up, use_noise
        self.activation, self.conv_clamp = activation, conv_clamp
        self.register_buffer('resample_filter', fir.setup_filter(list(resample_filter)))
        self.padding = kernel_size // 2
        self.act_gain = act.ACTIVATIONS[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        if use_noise:
            self.register_buffer('noise_const', torch.randn(resolution, resolution))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1):
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn(x.shape[0], 1, self.resolution, self.resolution, device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        x = conv.modulated_conv2d(x, self.weight, styles, noise=noise, up=self.up, padding=self.padding,
                                  resample_filter=self.resample_filter, flip_weight=(self.up == 1),
                                  fused_modconv=fused_modconv)
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return act.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=clamp)


class ToRGBLayer(torch.nn.Module):
    """[UPSTREAM] 1x1 modulated conv without demodulation, styles pre-scaled by 1/sqrt(Cin)."""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.weight_gain = 1 / math.sqrt(in_channels * kernel_size ** 2)

    def forward(self, x, w, fused_modconv=True):
        styles = self.affine(w) * self.weight_gain
        x = conv.modulated_conv2d(x, self.weight, styles, demodulate=False, fused_modconv=fused_modconv)
        return act.bias_act(x, self.bias.to(x.dtype), clamp=self.conv_clamp)


class SynthesisBlock(torch.nn.Module):
    """[UPSTREAM] one resolution: (const | conv0 up=2) -> conv1 -> torgb ('skip' architecture)."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=(1, 3, 3, 1), conv_clamp=None, use_fp16=False, fp16_channels_last=False,
                 **layer_kwargs):
        super().__init__()
        assert architecture == 'skip', 'FFHQ config-f networks are skip; orig/resnet are out of scope'
        self.in_channels, self.w_dim, self.resolution = in_channels, w_dim, resolution
        self.img_channels, self.is_last, self.architecture = img_channels, is_last, architecture
        self.use_fp16 = use_fp16
        self.channels_last = use_fp16 and fp16_channels_last
        self.register_buffer('resample_filter', fir.setup_filter(list(resample_filter)))
        self.num_conv = self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn(out_channels, resolution, resolution))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim, resolution, up=2,
                                        resample_filter=resample_filter, conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim, resolution, conv_clamp=conv_clamp,
                                    **layer_kwargs)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim, conv_clamp=conv_clamp)
        self.num_torgb += 1


class SynthesisNetwork(torch.nn.Module):
    """[UPSTREAM] blocks b4..b{img_resolution}; channels = min(channel_base // res, channel_max)."""

    def __init__(self, w_dim=512, img_resolution=1024, img_channels=3, channel_base=32768, channel_max=512,
                 num_fp16_res=0, **block_kwargs):
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        log2 = int(math.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, log2 + 1)]
        ch = {r: min(channel_base // r, channel_max) for r in self.block_resolutions}
        fp16_res = max(2 ** (log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for r in self.block_resolutions:
            blk = SynthesisBlock(ch[r // 2] if r > 4 else 0, ch[r], w_dim=w_dim, resolution=r,
                                 img_channels=img_channels, is_last=(r == img_resolution),
                                 use_fp16=(r >= fp16_res), **block_kwargs)
            self.num_ws += blk.num_conv + (blk.num_torgb if r == img_resolution else 0)
            setattr(self, f'b{r}', blk)


class MappingNetwork(torch.nn.Module):
    """[UPSTREAM MappingNetwork, c_dim = 0] z -> W+: second-moment normalisation, ``num_layers`` FullyConnectedLayers (lrelu, gain
    sqrt 2, lr multiplier 0.01), broadcast to ``num_ws`` rows, truncation ``w_avg.lerp(w, psi)``.  Constructor kwargs as
    ``legacy.py:129-136`` (mapping_layers 8, mapping_lrmul 0.01); attribute names ``fc{i}`` / ``w_avg`` as ``legacy.py:175-181``.
    Called once per seed set by generate_w.py:48-50."""

    def __init__(self, z_dim=512, w_dim=512, num_ws=18, num_layers=8, lr_multiplier=0.01):
        super().__init__()
        self.z_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, w_dim, num_ws, num_layers
        for i in range(num_layers):
            setattr(self, f'fc{i}', FullyConnectedLayer(z_dim if i == 0 else w_dim, w_dim, activation='lrelu', lr_multiplier=lr_multiplier))
        self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1, truncation_cutoff=None):
        x = z.to(self.w_avg.dtype)
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
    """Minimal stand-in for the unpickled G_ema: ``.synthesis`` is the hot path; ``.mapping`` (built AFTER the synthesis network, as
    upstream, so the synthesis weights of a seed do not depend on it) only when asked for."""

    def __init__(self, mapping=False, **synthesis_kwargs):
        super().__init__()
        self.synthesis = SynthesisNetwork(**synthesis_kwargs)
        self.z_dim, self.c_dim, self.w_dim = 512, 0, self.synthesis.w_dim
        if mapping:
            self.mapping = MappingNetwork(z_dim=self.z_dim, w_dim=self.w_dim, num_ws=self.synthesis.num_ws)


def generate_w(G, seeds, truncation_psi=1.0):
    """generate_w.py:46-50: one ``np.random.RandomState(seed).randn(1, z_dim)`` per seed -> ``G.mapping`` -> W+ [M, num_ws, 512]."""
    import numpy as np
    zs = torch.cat([torch.from_numpy(np.random.RandomState(seed).randn(1, G.z_dim)) for seed in seeds])
    p = next(G.mapping.parameters())
    with torch.no_grad():
        return G.mapping(zs.to(p.device, p.dtype), None, truncation_psi=truncation_psi)


def make_generator(img_resolution, seed=0, channel_base=32768, channel_max=512, conv_clamp=256,
                   noise_strength=0.1, torgb_scale=0.25, dtype=torch.float32, mapping=False):
    """Random-init config-f style generator (BASELINE.json config 1/2/4).

    Upstream init (weights ~N(0,1), bias 0, affine bias 1, noise_strength 0) with two documented
    deviations so the tests exercise more of the path: noise_strength := 0.1 (noise path live) and the
    ToRGB weights scaled by ``torgb_scale`` so the image mostly lies in [-1, 1] instead of
    saturating find_direction's clamp(0,255) on ~half the pixels.
    """
    gen = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.random.set_rng_state(gen.get_state())
    try:
        G = Generator(mapping=mapping, w_dim=512, img_resolution=img_resolution, img_channels=3, channel_base=channel_base,
                      channel_max=channel_max, num_fp16_res=0, conv_clamp=conv_clamp)
        if mapping:
            G.mapping.w_avg.copy_(0.1 * torch.randn(512))       # a trained network carries the running mean of W here
    finally:
        torch.random.set_rng_state(state)
    with torch.no_grad():
        for r in G.synthesis.block_resolutions:
            blk = getattr(G.synthesis, f'b{r}')
            blk.torgb.weight.mul_(torgb_scale)
            for name in ('conv0', 'conv1'):
                if hasattr(blk, name):
                    getattr(blk, name).noise_strength.fill_(noise_strength)
    return G.to(dtype).eval().requires_grad_(False)


# ------------------------------------------------------------------------------------------------
# S-space driver (reference utils.py)

def split_ws(G, ws):
    """utils.py:77-87: block k sees ws[:, w_idx : w_idx+num_conv+num_torgb]; w_idx += num_conv."""
    out, i = [], 0
    for r in G.synthesis.block_resolutions:
        blk = getattr(G.synthesis, f'b{r}')
        out.append(ws.to(torch.float32).narrow(1, i, blk.num_conv + blk.num_torgb))
        i += blk.num_conv
    return out


def _layers(blk):
    return [blk.conv1, blk.torgb] if blk.in_channels == 0 else [blk.conv0, blk.conv1, blk.torgb]


def get_temp_shapes(G):
    """utils.py:100-120: (C_conv0, C_conv1, C_torgb) per block, and every affine becomes Identity."""
    shapes = []
    for r in G.synthesis.block_resolutions:
        layers = _layers(getattr(G.synthesis, f'b{r}'))
        c = [l.affine.weight.shape[0] for l in layers]
        shapes.append((c[0], c[0], c[1]) if len(c) == 2 else tuple(c))
        for l in layers:
            l.affine = torch.nn.Identity()
    return shapes


def get_styles(G, ws, block_ws):
    """utils.py:123-158: S [M,26,512] zero padded, row j = affine_j(w); affines replaced by Identity."""
    styles = torch.zeros(ws.shape[0], N_STYLE_ROWS, STYLE_WIDTH, dtype=ws.dtype, device=ws.device)
    shapes, row = [], 0
    with torch.no_grad():
        for r, cur in zip(G.synthesis.block_resolutions, block_ws):
            layers = _layers(getattr(G.synthesis, f'b{r}'))
            c = [l.affine.weight.shape[0] for l in layers]
            shapes.append((c[0], c[0], c[1]) if len(c) == 2 else tuple(c))
            for j, l in enumerate(layers):
                styles[:, row, :c[j]] = l.affine(cur[:, j, :].to(ws.dtype))
                l.affine = torch.nn.Identity()
                row += 1
    return styles, shapes


def block_forward(blk, x, img, s, shapes, noise_mode='const'):
    """utils.py:13-53 for the fp32 'skip' case (fused_modconv = not training, :18-20)."""
    fused = not blk.training
    rows = iter(s.unbind(dim=1))
    if blk.in_channels == 0:
        x = blk.const.to(s.dtype).unsqueeze(0).repeat(s.shape[0], 1, 1, 1)
        x = blk.conv1(x, next(rows)[..., :shapes[0]], fused_modconv=fused, noise_mode=noise_mode)
    else:
        x = blk.conv0(x, next(rows)[..., :shapes[0]], fused_modconv=fused, noise_mode=noise_mode)
        x = blk.conv1(x, next(rows)[..., :shapes[1]], fused_modconv=fused, noise_mode=noise_mode)
    if img is not None:
        img = fir.upsample2d(img, blk.resample_filter)
    y = blk.torgb(x, next(rows)[..., :shapes[2]], fused_modconv=fused)
    img = y if img is None else img + y
    return x, img


def generate_image(G, until_k, styles, temp_shapes, noise_mode='const'):
    """utils.py:161-216 without blending: returns (per-block feature maps, running skip image)."""
    x = img = None
    xs, row = [], 0
    for k, r in enumerate(G.synthesis.block_resolutions):
        if k > until_k:
            continue
        n = 2 if r == 4 else 3
        x, img = block_forward(getattr(G.synthesis, f'b{r}'), x, img, styles[:, row:row + n, :], temp_shapes[k],
                               noise_mode)
        row += n
        xs.append(x)
    return xs, img
