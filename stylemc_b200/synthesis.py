"""Fused S-space synthesis engine: the B200 execution plan behind ``utils.generate_image``.

The reference walks ``G.synthesis.b{res}`` and calls, per layer, modulated_conv2d -> conv2d_resample -> (upfirdn2d) ->
bias_act on NCHW fp32 tensors (utils.py:13-53,161-216; [UPSTREAM] training/networks.py).  Here the same arithmetic runs
as a short chain of hand-written kernels over NHWC fp16 activations that never leave the GPU:

  conv1 (3x3)        one tcgen05 implicit GEMM; demodulation, noise, bias, lrelu*sqrt2 and clamp in its epilogue
  conv0 (3x3, up 2)  four parity implicit GEMMs (polyphase transposed conv, no multiply-by-zero) -> one kernel doing the
                     4x4 FIR + noise + bias + lrelu + clamp and the next layer's style multiply
  torgb + skip       one kernel: 1x1 modulated conv to RGB, clamp, upsample2d of the running image, add, and the style
                     multiply for the next block's conv0
  backward           act_bwd (activation/clamp mask, ToRGB branch, demodulation, style-gradient reductions) -> dgrad
                     implicit GEMM (-> transposed FIR for conv0); only the S rows that are trainable get reductions, and no
                     per-sample weight gradient is ever formed (SURVEY.md section 8a style-gradient algebra).

Styles are applied to the activations (the reference's non-fused formulation, identical maths -- oracle/pin_reference.py
checks fused vs non-fused), so the frozen weights are shared by the whole batch.
"""
import ctypes
import math
import os

import torch

from . import _lib, gemm
from .ops import upfirdn2d

N_STYLE_ROWS = 26        # find_direction.py:38
STYLE_WIDTH = 512        # utils.py:125
LRELU_ALPHA = 0.2
LRELU_GAIN = math.sqrt(2)


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class _Layer:
    """Device-resident frozen parameters of one SynthesisLayer in the layouts the kernels want."""

    def __init__(self, mod, device, up):
        w = _f32(mod.weight, device)                                   # [O, I, 3, 3]
        self.cout, self.cin = w.shape[:2]
        self.up = up
        self.resolution = int(mod.resolution)
        # one kernel: hi/lo planes of the tap matrices, forward (row t*O + o, col i) and dgrad (row t*I + i, col o), and q = sum_t w^2
        self.B_fwd, self.B_bwd, self.q, _ = gemm.prepare_weights(w, two=True, fwd=True, bwd=True, q=True)
        self.bias = _f32(mod.bias, device)
        strength = float(mod.noise_strength) if getattr(mod, 'use_noise', True) else 0.0
        self.noise_const = (_f32(mod.noise_const, device) * strength).contiguous() if getattr(mod, 'use_noise', True) else None
        self.noise_strength = strength
        self.clamp = float(mod.conv_clamp) if mod.conv_clamp is not None else -1.0
        self.gain = float(getattr(mod, 'act_gain', LRELU_GAIN))
        assert getattr(mod, 'activation', 'lrelu') == 'lrelu', 'the fused engine implements the lrelu synthesis layers'


class _ToRGB:
    def __init__(self, mod, device):
        w = _f32(mod.weight, device)                                   # [3, C, 1, 1]
        assert w.shape[0] == 3 and w.shape[2:] == (1, 1)
        self.cin = w.shape[1]
        self.w = w.reshape(3, self.cin).contiguous()
        self.bias = _f32(mod.bias, device)
        self.wgain = float(mod.weight_gain)
        self.clamp = float(mod.conv_clamp) if mod.conv_clamp is not None else -1.0


class _Block:
    def __init__(self, blk, device):
        self.resolution = int(blk.resolution)
        self.const = _f32(blk.const, device) if int(blk.in_channels) == 0 else None
        self.conv0 = _Layer(blk.conv0, device, up=2) if self.const is None else None
        self.conv1 = _Layer(blk.conv1, device, up=1)
        self.torgb = _ToRGB(blk.torgb, device)
        assert getattr(blk, 'architecture', 'skip') == 'skip'


class SavedForward:
    """What one forward pass keeps for the backward pass (fp16 layer outputs, demodulation coefficients)."""

    def __init__(self):
        self.y0, self.y1, self.d0, self.d1 = {}, {}, {}, {}
        self.rgb_pass = {}        # block -> [N, 3, H, W] uint8, 1 where the ToRGB clamp passes the gradient (fused-ToRGB blocks only)
        self.styles = None
        self.until_k = None


class SynthesisEngine:
    """Execution plan for one frozen ``G.synthesis`` on one GPU.

    precision: 'x1' fp16 operands (fast); 'x3' split-fp16 operands everywhere; 'mixed' = 'x3' for blocks up to ``x3_max_res``
    and 'x1' above; 'x3p' = 'x3' with promoted accumulation (the K loop is drained into fp32 registers every 512 elements),
    the mode that reproduces an fp32 reference to fp32 accuracy -- see DESIGN.md "Numerics".
    """

    def __init__(self, G, device='cuda', precision='mixed', x3_max_res=64):
        syn = G.synthesis if hasattr(G, 'synthesis') else G
        self.device = torch.device(device)
        self.block_resolutions = list(syn.block_resolutions)
        self.blocks = [_Block(getattr(syn, f'b{r}'), self.device) for r in self.block_resolutions]
        self.img_resolution = int(syn.img_resolution)
        f = upfirdn2d.setup_filter([1, 3, 3, 1], device=self.device)
        blk0 = getattr(syn, f'b{self.block_resolutions[0]}')
        if hasattr(blk0, 'resample_filter'):
            f = _f32(blk0.resample_filter, self.device)
        assert f.shape == (4, 4), 'the fused engine implements the 4x4 [1,3,3,1] resample filter'
        self.filter = f
        self.fk4 = (f.flip([0, 1]) * 4.0).contiguous()                 # flipped taps * gain (up=2 -> gain 4)
        # separable? (the [1,3,3,1] filter is): host copy of the factors fk4[fy][fx] = fy[fy] * fx[fx] for the marching FIR kernels
        fh = self.fk4.double().cpu()
        rs, cs, tot = fh.sum(1), fh.sum(0), fh.sum()
        self.fsep = None
        if abs(float(tot)) > 1e-20 and float((torch.outer(rs, cs) / tot - fh).abs().max()) <= 1e-6 * float(fh.abs().max()):
            self.fsep = (ctypes.c_float * 8)(*[float(v) for v in (rs / tot)], *[float(v) for v in cs])
        if precision not in ('x1', 'x3', 'mixed', 'x3p'):
            raise ValueError(precision)
        self.acc_k = 512 if precision == 'x3p' else 0
        self.acc_k_lowres, self.acc_k_lowres_max = 64, 16
        self.fuse_torgb = os.environ.get('STYLEMC_HCONV') != '0'      # the fused epilogue exists in hconv.cu only
        self.group_parities = os.environ.get('STYLEMC_CONV0_GROUP') != '0'  # conv0: the four parity GEMMs as one problem-group launch
        self.fuse_act_bwd = os.environ.get('STYLEMC_FUSE_ACT_BWD') != '0'   # also needs fuse_torgb (both live in hconv.cu)
        self.fuse_rgb_wide = os.environ.get('STYLEMC_FUSE_RGB_WIDE') != '0'  # fused ToRGB also for layers of several N tiles (cout > 128)
        # backward pass: hi + lo planes for the activation GRADIENTS too (3 MMAs per product) instead of a hi plane (2 MMAs); measured
        # identical style gradients to the 4th digit (tests/diag/diag_grad_planes.py), so off by default
        self.grad_lo = os.environ.get('STYLEMC_GRAD_LO', '0') != '0'
        self.bwd_prec = os.environ.get('STYLEMC_BWD_PREC', 'x2')           # 'x1' (weights hi plane only) is a diagnostic: diag_grad_planes.py
        self.precision, self.x3_max_res = ('x3' if precision == 'x3p' else precision), x3_max_res
        # style row of (conv0, conv1, torgb) per block (utils.py:169-185)
        self.rows, r = [], 0
        for b in self.blocks:
            if b.conv0 is None:
                self.rows.append((None, r, r + 1))
                r += 2
            else:
                self.rows.append((r, r + 1, r + 2))
                r += 3

    # ---- helpers -------------------------------------------------------------------------------
    def _prec(self, res):
        if self.precision == 'mixed':
            return 'x3' if res <= self.x3_max_res else 'x1'
        return self.precision

    def _acc_k(self, res):
        """K elements per main accumulation chain of a FORWARD conv whose output is res x res.  A leaky-ReLU unit whose pre-activation
        lies within the forward rounding error of zero takes the wrong slope in the backward pass; at low resolution one such unit
        moves the style gradient by ~1e-3, and the low-resolution layers cost next to nothing, so they get
        chains of 4 MMAs (64 K elements): forward error ~1e-7, like the fp32 reference."""
        if self.acc_k and res <= self.acc_k_lowres_max:
            return self.acc_k_lowres
        return self.acc_k

    @staticmethod
    def _hconv_shape(res, cin, cout):
        """Shapes the halo-tile kernel (csrc/hconv.cu) takes in its default routing: the fused epilogues exist there only."""
        return res >= 32 and cin % 32 == 0 and cout % 32 == 0

    def _fsep_ptr(self):
        return ctypes.addressof(self.fsep) if self.fsep is not None else None

    @staticmethod
    def _srow(styles, row):
        """(device pointer of styles[0, row, 0], element stride between samples)."""
        return styles.data_ptr() + row * styles.stride(1) * 4, styles.stride(0)

    def _demod(self, layer, styles, row, n):
        d = torch.empty([n, layer.cout], dtype=torch.float32, device=self.device)
        sp, ss = self._srow(styles, row)
        _lib.call('smc_demod_coefs', _lib.ptr(layer.q), sp, ss, _lib.ptr(d), n, layer.cin, layer.cout, _lib.stream())
        return d

    @staticmethod
    def _masked_grgb(g_img, rgb_pass, gscale):
        """g_img * rgb_pass * gscale in one kernel: the incoming image gradient under the ToRGB clamp mask of the forward pass, loss-scaled."""
        out = torch.empty_like(g_img)
        _lib.call('smc_mask_scale', _lib.ptr(g_img), _lib.ptr(rgb_pass), _lib.ptr(gscale), _lib.ptr(out), g_img.numel(), _lib.stream())
        return out

    def _planes(self, n, h, w, c, two):
        return torch.empty([2 if two else 1, n, h, w, c], dtype=torch.float16, device=self.device)

    def _noise(self, layer, noise_mode, n):
        if noise_mode == 'none' or layer.noise_const is None:
            return None
        if noise_mode == 'const':
            return layer.noise_const
        raise RuntimeError("the fused engine implements noise_mode 'const' and 'none' (find_direction.py:207 default 'const'); "
                           "'random' draws per-sample noise and is not reproducible")

    def _conv1(self, L, xs, d, noise, n, res, prec, want_lo):
        """3x3 modulated conv + noise + bias + lrelu + clamp; returns y planes [P, n, res, res, cout]."""
        y = self._planes(n, res, res, L.cout, want_lo)
        gemm.igemm(xs.reshape(-1, res, res, L.cin), L.B_fwd, n, res, res, L.cout, gemm.TAPS_3X3, precision=prec, acc_chunk_k=self._acc_k(res),
                   a_plane_stride_imgs=n, b_rows_per_tap=9 * L.cout, row_scale=d, bias=L.bias, noise=noise,
                   noise_strides=(res, 1), act=1, alpha=LRELU_ALPHA, gain=L.gain, clamp=L.clamp,
                   out_hi=y[0], out_lo=y[1] if want_lo else None)
        return y

    @staticmethod
    def _rgb_parts(L):
        """N tiles of the halo-tile kernel for this layer's forward conv (csrc/hconv.cu picks 128 / 64 / 32 by divisibility): the fused ToRGB
        keeps one partial-sum image per N tile (two threads add into each value of it, so the atomic accumulation is order-independent),
        and smc_img_finish adds the images in index order: deterministic whatever the channel count."""
        bn = 128 if L.cout % 128 == 0 else (64 if L.cout % 64 == 0 else 32)
        return L.cout // bn

    def _fusable(self, L, res):
        """conv1 layers that run on the halo-tile kernel (csrc/hconv.cu) take ToRGB and the next style multiply in their epilogue: every block
        from 32 px up (round 1 / 2: only the single-N-tile layers from 256 px up; the 512 / 256-channel blocks ran smc_torgb, a separate
        pass over y, and had no saved ToRGB clamp mask for the fused activation backward).  STYLEMC_FUSE_RGB_WIDE=0 restores that."""
        return res >= 32 and L.cin % 32 == 0 and L.cout % 32 == 0 and (self.fuse_rgb_wide or L.cout <= 128)

    def _conv1_fused(self, L, T, xs, d, noise, styles, rt, row_next, n, res, prec, two, keep_y, next_two, y_full=True):
        """conv1 with everything that consumes its output fused into the GEMM epilogue: the saved activation y (hi/lo, only when
        it is needed), xs_next = y * styles[:, row_next] for the next block's conv0, and the ToRGB 1x1 modulated conv accumulated
        into zeroed fp32 images, one per N tile of the kernel (summed and finished by smc_img_finish).  Returns (y or None, xs_next or None,
        rgb accumulators [parts, n, 3, res, res])."""
        y = self._planes(n, res, res, L.cout, two and y_full) if keep_y else None
        xn = self._planes(n, res, res, L.cout, next_two) if row_next is not None else None
        post = styles[:, row_next, :L.cout].contiguous() if row_next is not None else None
        rgb_w = ((styles[:, rt, :L.cout] * T.wgain).unsqueeze(1) * T.w.unsqueeze(0)).contiguous()        # [n, 3, C]
        parts = self._rgb_parts(L)
        acc = torch.zeros([parts, n, 3, res, res], dtype=torch.float32, device=self.device)
        gemm.igemm(xs.reshape(-1, res, res, L.cin), L.B_fwd, n, res, res, L.cout, gemm.TAPS_3X3, precision=prec, acc_chunk_k=self._acc_k(res),
                   a_plane_stride_imgs=n, b_rows_per_tap=9 * L.cout, row_scale=d, bias=L.bias, noise=noise,
                   noise_strides=(res, 1), act=1, alpha=LRELU_ALPHA, gain=L.gain, clamp=L.clamp,
                   out_raw=y[0] if keep_y else None, out_raw_lo=y[1] if (keep_y and y.shape[0] == 2) else None,
                   post_scale=post, out_hi=xn[0] if xn is not None else None, out_lo=xn[1] if (xn is not None and next_two) else None,
                   rgb_w=rgb_w, rgb_acc=acc[0], rgb_part_stride=acc.stride(0) if parts > 1 else 0)
        return y, xn, acc

    def _conv0(self, L, xs, d, noise, styles, row_next, n, hin, prec, want_lo, save_lo=False, save_full=True):
        """3x3 transposed stride-2 modulated conv + 4x4 FIR + noise + bias + lrelu + clamp.
        Returns (y planes [P, n, 2h, 2h, cout] (lo only when saving for an x3 backward), xs_next planes = y * styles[:, row_next])."""
        x3 = prec == 'x3'
        planes = torch.empty([4, n, hin + 1, hin + 1, L.cout], dtype=torch.float32 if x3 else torch.float16, device=self.device)
        if self.group_parities:
            # one launch for the four output parities (problem group): every input tile comes from DRAM once
            taps, problems = [], []
            for q, (r, c) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                t = gemm.up2_parity_taps(r, c)
                taps += t
                problems.append((len(t), q * planes[0].numel()))
            kw = dict(out_f32=planes[0]) if x3 else dict(out_raw=planes[0])
            gemm.igemm(xs.reshape(-1, hin, hin, L.cin), L.B_fwd, n, hin + 1, hin + 1, L.cout, taps, precision=prec, acc_chunk_k=self._acc_k(2 * hin),
                       a_plane_stride_imgs=n, b_rows_per_tap=9 * L.cout, row_scale=d, problems=problems, **kw)
        else:
            for r in (0, 1):
                for c in (0, 1):
                    kw = dict(out_f32=planes[r * 2 + c]) if x3 else dict(out_raw=planes[r * 2 + c])
                    gemm.igemm(xs.reshape(-1, hin, hin, L.cin), L.B_fwd, n, hin + 1, hin + 1, L.cout, gemm.up2_parity_taps(r, c),
                               precision=prec, acc_chunk_k=self._acc_k(2 * hin), a_plane_stride_imgs=n, b_rows_per_tap=9 * L.cout, row_scale=d, **kw)
        res = 2 * hin
        # the raw activation is only needed by the backward pass; its lo plane only where a style-gradient reduction reads it
        y = self._planes(n, res, res, L.cout, x3 and save_full) if save_lo else None
        xn = self._planes(n, res, res, L.cout, want_lo)
        sp, ss = self._srow(styles, row_next)
        _lib.call('smc_fir_act', _lib.ptr(planes), 0 if x3 else 1, n, hin, hin, L.cout, _lib.ptr(self.fk4), self._fsep_ptr(), _lib.ptr(noise),
                  _lib.ptr(L.bias), LRELU_ALPHA, L.gain, L.clamp, sp, ss, _lib.ptr(y[0]) if y is not None else None,
                  _lib.ptr(y[1]) if (y is not None and y.shape[0] == 2) else None,
                  _lib.ptr(xn[0]), _lib.ptr(xn[1]) if want_lo else None, _lib.stream())
        return y, xn

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, styles, until_k=100, noise_mode='const', save=False, want_xs=False, grad_rows=None):
        """styles [N, 26, 512] fp32 CUDA -> (xs list or None, img [N, 3, R, R] fp32, SavedForward or None).

        Mirrors utils.generate_image (utils.py:161-216) without the blending branches.
        grad_rows (with save=True): the style rows whose gradient ``backward`` will be asked for.  Layers that feed none of their
        style-gradient reductions keep only the hi plane of their saved activation (the backward pass needs just its sign and
        clamp mask there); None keeps every plane."""
        _lib.require_cuda(styles, 'styles')
        if styles.ndim != 3 or styles.shape[1] < self.rows[min(until_k, len(self.blocks) - 1)][2] + 1:
            raise RuntimeError(f'styles must be [N, >= {self.rows[-1][2] + 1}, C], got {tuple(styles.shape)}')
        styles = styles.float().contiguous()
        n = styles.shape[0]
        saved = SavedForward() if save else None
        xs_list = [] if want_xs else None
        img = None
        xs = None                      # input planes of the next conv (activation times that conv's styles)
        with torch.cuda.device(self.device):
            last = min(until_k, len(self.blocks) - 1)
            for k, blk in enumerate(self.blocks[:last + 1]):
                res = blk.resolution
                r0, r1, rt = self.rows[k]
                prec = self._prec(res)
                two = prec == 'x3'
                L1 = blk.conv1
                if blk.const is not None:
                    c = blk.const.shape[0]
                    xs = self._planes(n, res, res, c, two)
                    sp, ss = self._srow(styles, r1)
                    _lib.call('smc_pack_nhwc', _lib.ptr(blk.const), 0, sp, ss, _lib.ptr(xs[0]), _lib.ptr(xs[1]) if two else None,
                              n, c, res * res, c, _lib.stream())
                else:
                    L0 = blk.conv0
                    d0 = self._demod(L0, styles, r0, n)
                    y0, xs = self._conv0(L0, xs, d0, self._noise(L0, noise_mode, n), styles, r1, n, res // 2, prec, two, save_lo=save,
                                         save_full=grad_rows is None or r0 in grad_rows or r1 in grad_rows)
                    if save:
                        saved.y0[k], saved.d0[k] = y0, d0
                d1 = self._demod(L1, styles, r1, n)
                T = blk.torgb
                has_next = k < last
                nprec_two = has_next and self._prec(self.blocks[k + 1].resolution) == 'x3'
                if self.fuse_torgb and self._fusable(L1, res):
                    y1, xs_next, new_img = self._conv1_fused(L1, T, xs, d1, self._noise(L1, noise_mode, n), styles, rt,
                                                             self.rows[k + 1][0] if has_next else None, n, res, prec, two,
                                                             keep_y=save or want_xs, next_two=nprec_two,
                                                             y_full=want_xs or grad_rows is None or r1 in grad_rows or
                                                             (has_next and self.rows[k + 1][0] in grad_rows))
                    rgb_pass = torch.empty([n, 3, res, res], dtype=torch.uint8, device=self.device) if save else None
                    _lib.call('smc_img_finish', _lib.ptr(new_img), _lib.ptr(img), _lib.ptr(T.bias), T.clamp, _lib.ptr(self.fk4), n, res, res,
                              _lib.ptr(rgb_pass), new_img.shape[0], new_img.stride(0), _lib.stream())
                    new_img = new_img[0]                      # the partial sums of the N tiles were added into plane 0
                    if rgb_pass is not None:
                        saved.rgb_pass[k] = rgb_pass
                    xs = xs_next
                else:
                    y1 = self._conv1(L1, xs, d1, self._noise(L1, noise_mode, n), n, res, prec, two)
                    # ToRGB + skip + style multiply for the next block's conv0
                    new_img = torch.empty([n, 3, res, res], dtype=torch.float32, device=self.device)
                    if has_next:
                        xs = self._planes(n, res, res, L1.cout, nprec_two)
                        snp, sns = self._srow(styles, self.rows[k + 1][0])
                    stp, sts = self._srow(styles, rt)
                    _lib.call('smc_torgb', _lib.ptr(y1[0]), _lib.ptr(y1[1]) if two else None, n, res, res, L1.cout, _lib.ptr(T.w), stp, sts,
                              T.wgain, _lib.ptr(T.bias), T.clamp, _lib.ptr(img), _lib.ptr(self.fk4), _lib.ptr(new_img),
                              snp if has_next else None, sns if has_next else 0, _lib.ptr(xs[0]) if has_next else None,
                              _lib.ptr(xs[1]) if (has_next and xs.shape[0] == 2) else None, _lib.stream())
                if save:
                    saved.y1[k], saved.d1[k] = y1, d1
                img = new_img
                if want_xs:
                    out = torch.empty([n, L1.cout, res, res], dtype=torch.float32, device=self.device)
                    _lib.call('smc_unpack_nchw', _lib.ptr(y1[0]), 1, _lib.ptr(out), None, n, L1.cout, res * res, L1.cout, _lib.stream())
                    if y1.shape[0] == 2:   # add the lo plane so xs carries the full precision that was computed
                        lo = torch.empty_like(out)
                        _lib.call('smc_unpack_nchw', _lib.ptr(y1[1]), 1, _lib.ptr(lo), None, n, L1.cout, res * res, L1.cout, _lib.stream())
                        out += lo
                    xs_list.append(out)
        if save:
            saved.styles, saved.until_k = styles, last
            saved.grad_rows = None if grad_rows is None else set(grad_rows)
        return xs_list, img, saved

    # ---- backward ------------------------------------------------------------------------------
    def backward(self, saved, g_img, trainable_rows, noise_mode='const', grad_scale_target=256.0, per_sample=False, grad_lo=None):
        """Gradient of a scalar loss w.r.t. the trainable S rows, summed over the batch.

        saved: SavedForward of the pass that produced img; g_img = dL/dimg [N, 3, R, R] fp32.
        Returns grad [len(trainable_rows), 512] fp32 (zero beyond each layer's channel count), i.e. exactly
        ``delta.grad`` of find_direction.py:336 for delta broadcast over the batch (:307-308).  ``per_sample=True`` returns
        ``(grad, grad_samples [N, len(trainable_rows), 512])``: the gradient w.r.t. each image's own S rows (the latent mapper's delta differs
        per image, train_latent_mapper.py:155-158).  ``grad_lo`` overrides the engine's choice of gradient planes (hi only / hi + lo) for this
        pass: a per-sample gradient has no batch sum for the rounding of a single fp16 plane to average out in."""
        styles, last = saved.styles, saved.until_k
        use_grad_lo = self.grad_lo if grad_lo is None else bool(grad_lo)
        n = styles.shape[0]
        dev = self.device
        trainable_rows = list(trainable_rows)
        owner = {}
        for k in range(last + 1):
            r0, r1, rt = self.rows[k]
            if r0 is not None:
                owner[r0] = (k, 0)
            owner[r1] = (k, 1)
        for r in trainable_rows:
            if r not in owner:
                raise RuntimeError(f'style row {r} is not a conv layer of the blocks that ran (ToRGB rows are not trainable here)')
        lowest_k = min(owner[r][0] for r in trainable_rows)
        want = set(trainable_rows)
        if getattr(saved, 'grad_rows', None) is not None and not want <= saved.grad_rows:
            raise RuntimeError(f'the forward pass saved activations for style rows {sorted(saved.grad_rows)} only; asked for {sorted(want)}')
        grad = torch.zeros([len(trainable_rows), STYLE_WIDTH], dtype=torch.float32, device=dev)
        grad_samples = torch.zeros([n, len(trainable_rows), STYLE_WIDTH], dtype=torch.float32, device=dev) if per_sample else None
        g_img = g_img.float().contiguous()
        with torch.cuda.device(dev):
            amax = torch.zeros(1, dtype=torch.int32, device=dev)
            gscale = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.call('smc_grad_scale', _lib.ptr(g_img), g_img.numel(), grad_scale_target, _lib.ptr(amax), _lib.ptr(gscale), _lib.stream())
            acc = {}     # row -> (T1, R)

            def bufs(row, cin, cout):
                if row not in acc:
                    acc[row] = (torch.zeros([n, cin], dtype=torch.float32, device=dev), torch.zeros([n, cout], dtype=torch.float32, device=dev))
                return acc[row]

            g_up, up_row, up_f32 = None, None, False   # gradient w.r.t. the modulated input of the consumer conv above, its style row, dtype
            gd1_fused = None                           # conv1's gd planes already produced by the epilogue of the dgrad GEMM above
            for k in range(last, -1, -1):
                blk = self.blocks[k]
                res = blk.resolution
                r0, r1, rt = self.rows[k]
                L1, T = blk.conv1, blk.torgb
                y1, d1 = saved.y1[k], saved.d1[k]
                prec = self._prec(res)
                two = prec == 'x3'             # split-precision backward for the blocks whose forward was split-precision
                # gradient operands: the weights keep both planes; the activation gradients (loss-scaled, zero-mean rounding that averages
                # out in the style-gradient sums) carry a lo plane only when asked to (2 MMAs per product, not 3)
                gtwo = two and use_grad_lo
                gprec = prec if (gtwo or not two) else self.bwd_prec
                stop_here = (k < lowest_k)       # below the lowest trainable block only T1 of the consumer is needed
                # ---- conv1 output: consumers are ToRGB (g_img) and the next block's conv0 (g_up)
                t1 = bufs(up_row, L1.cout, self.blocks[k + 1].conv0.cout)[0] if (g_up is not None and up_row in want) else None
                rr = bufs(r1, L1.cin, L1.cout)[1] if r1 in want else None
                need_gd = not stop_here
                if gd1_fused is not None:
                    gd1, gd1_fused = gd1_fused, None
                else:
                    if not need_gd and t1 is None:
                        break
                    gd1 = self._planes(n, res, res, L1.cout, gtwo) if need_gd else None
                    sp, ss = self._srow(styles, up_row) if g_up is not None else (None, 0)
                    stp, sts = self._srow(styles, rt)
                    noise1 = self._noise(L1, noise_mode, n)
                    gi, rgb_clamp, gs = (g_img if need_gd else None), T.clamp, gscale
                    if need_gd and k in saved.rgb_pass and self.fuse_act_bwd and (self.fuse_rgb_wide or (g_up is None and t1 is None and rr is None)):
                        # the ToRGB clamp mask saved by the forward pass and the loss scale go into the incoming gradient, so the kernel does
                        # not recompute the ToRGB output (a second pass over y1); at the top block, where no reduction is wanted either, it
                        # does not read the lo plane of y1 (smc_act_bwd picks act_bwd_rgb_kernel)
                        gi, rgb_clamp, gs = self._masked_grgb(g_img, saved.rgb_pass[k], gscale), -1.0, None
                    _lib.call('smc_act_bwd', _lib.ptr(y1[0]), _lib.ptr(y1[1]) if y1.shape[0] == 2 else None, n, res, res, L1.cout,
                              _lib.ptr(g_up), int(up_f32), sp, ss, _lib.ptr(gi), _lib.ptr(T.w), stp, sts, T.wgain,
                              _lib.ptr(T.bias), rgb_clamp, _lib.ptr(gs), _lib.ptr(d1), _lib.ptr(noise1), _lib.ptr(L1.bias), LRELU_ALPHA,
                              L1.gain, L1.clamp, _lib.ptr(gd1[0]) if need_gd else None, _lib.ptr(gd1[1]) if (need_gd and gtwo) else None,
                              _lib.ptr(t1), _lib.ptr(rr), _lib.stream())
                if stop_here:
                    break
                # ---- dgrad conv1 -> gradient w.r.t. (y0 * s1) (or const * s1 for b4)
                if blk.conv0 is None:
                    if r1 in want:
                        raise RuntimeError('trainable style rows in b4 are not implemented (reference trains b8..b64 only)')
                    break
                L0 = blk.conv0
                y0, d0 = saved.y0[k], saved.d0[k]
                t1 = bufs(r1, L1.cin, L1.cout)[0] if r1 in want else None
                rr = bufs(r0, L0.cin, L0.cout)[1] if r0 in want else None
                gd0 = self._planes(n, res, res, L0.cout, gtwo)
                if self.fuse_torgb and self.fuse_act_bwd and t1 is None and rr is None and self._hconv_shape(res, L1.cout, L1.cin):
                    # no style-gradient reduction wanted from this layer: the activation backward of conv0 (slope and clamp mask of the
                    # saved y0, conv1's style, conv0's demodulation) is the epilogue of conv1's dgrad GEMM -- no fp32 round trip
                    post = (styles[:, r1, :L1.cin] * d0).contiguous()
                    gemm.igemm(gd1.reshape(-1, res, res, L1.cout), L1.B_bwd, n, res, res, L1.cin, gemm.TAPS_3X3_DGRAD, precision=gprec,
                               acc_chunk_k=self.acc_k, a_plane_stride_imgs=n, b_rows_per_tap=9 * L1.cin, post_scale=post, alpha=LRELU_ALPHA,
                               gain=L0.gain, clamp=L0.clamp, mask_y=y0[0], mask_y_lo=y0[1] if y0.shape[0] == 2 else None,
                               out_hi=gd0[0], out_lo=gd0[1] if gtwo else None)
                else:
                    gx1 = torch.empty([n, res, res, L1.cin], dtype=torch.float32 if two else torch.float16, device=dev)
                    gemm.igemm(gd1.reshape(-1, res, res, L1.cout), L1.B_bwd, n, res, res, L1.cin, gemm.TAPS_3X3_DGRAD, precision=gprec,
                               acc_chunk_k=self.acc_k, a_plane_stride_imgs=n, b_rows_per_tap=9 * L1.cin,
                               **(dict(out_f32=gx1) if two else dict(out_raw=gx1)))
                    sp, ss = self._srow(styles, r1)
                    noise0 = self._noise(L0, noise_mode, n)
                    _lib.call('smc_act_bwd', _lib.ptr(y0[0]), _lib.ptr(y0[1]) if y0.shape[0] == 2 else None, n, res, res, L0.cout, _lib.ptr(gx1),
                              int(two), sp, ss, None, None, None, 0, 0.0, None, -1.0, _lib.ptr(gscale), _lib.ptr(d0), _lib.ptr(noise0),
                              _lib.ptr(L0.bias), LRELU_ALPHA, L0.gain, L0.clamp, _lib.ptr(gd0[0]), _lib.ptr(gd0[1]) if gtwo else None,
                              _lib.ptr(t1), _lib.ptr(rr), _lib.stream())
                hin = res // 2
                gp = torch.empty([2 if gtwo else 1, 4 * n, hin + 1, hin + 1, L0.cout], dtype=torch.float16, device=dev)
                _lib.call('smc_fir_bwd', _lib.ptr(gd0[0]), _lib.ptr(gd0[1]) if gtwo else None, n, hin, hin, L0.cout, _lib.ptr(self.fk4), self._fsep_ptr(),
                          _lib.ptr(gp[0]), _lib.ptr(gp[1]) if gtwo else None, _lib.stream())
                # ---- skip image: transpose of upsample2d (upfirdn2d.py:245-264)
                if k > 0:
                    g_img = upfirdn2d.upfirdn2d(g_img, self.filter, down=2, padding=[1, 1, 1, 1], flip_filter=True, gain=4)
                # ---- dgrad conv0 -> gradient w.r.t. (y1 of the block below * s0)
                Lp, Tp = self.blocks[k - 1].conv1, self.blocks[k - 1].torgb
                pr1, prt = self.rows[k - 1][1], self.rows[k - 1][2]
                two_p = self._prec(self.blocks[k - 1].resolution) == 'x3'
                if (self.fuse_torgb and self.fuse_act_bwd and (k - 1) in saved.rgb_pass and k - 1 >= lowest_k and r0 not in want and pr1 not in want
                        and two_p == two and self._hconv_shape(hin, L0.cout, L0.cin)):
                    # the block below needs no style-gradient reduction either: its conv1 activation backward, ToRGB branch included
                    # (g_rgb masked by the clamp mask saved in the forward pass), is the epilogue of this dgrad GEMM
                    y1p, d1p = saved.y1[k - 1], saved.d1[k - 1]
                    grgb = self._masked_grgb(g_img, saved.rgb_pass[k - 1], gscale)
                    post = (styles[:, r0, :L0.cin] * d1p).contiguous()
                    rgbw = ((styles[:, prt, :Lp.cout] * Tp.wgain * d1p).unsqueeze(1) * Tp.w.unsqueeze(0)).contiguous()      # [n, 3, C]
                    gd1_fused = self._planes(n, hin, hin, Lp.cout, gtwo)
                    gemm.igemm(gp.reshape(-1, hin + 1, hin + 1, L0.cout), L0.B_bwd, n, hin, hin, L0.cin, gemm.up2_dgrad_taps(n), precision=gprec,
                               acc_chunk_k=self.acc_k, a_plane_stride_imgs=4 * n, b_rows_per_tap=9 * L0.cin, post_scale=post, alpha=LRELU_ALPHA,
                               gain=Lp.gain, clamp=Lp.clamp, mask_y=y1p[0], mask_y_lo=y1p[1] if y1p.shape[0] == 2 else None, mask_grgb=grgb,
                               rgb_w=rgbw, out_hi=gd1_fused[0], out_lo=gd1_fused[1] if gtwo else None)
                    g_up = None
                else:
                    g_up = torch.empty([n, hin, hin, L0.cin], dtype=torch.float32 if two else torch.float16, device=dev)
                    gemm.igemm(gp.reshape(-1, hin + 1, hin + 1, L0.cout), L0.B_bwd, n, hin, hin, L0.cin, gemm.up2_dgrad_taps(n), precision=gprec,
                               acc_chunk_k=self.acc_k, a_plane_stride_imgs=4 * n, b_rows_per_tap=9 * L0.cin,
                               **(dict(out_f32=g_up) if two else dict(out_raw=g_up)))
                up_row, up_f32 = r0, two
            for i, row in enumerate(trainable_rows):
                k, which = owner[row]
                L = self.blocks[k].conv1 if which == 1 else self.blocks[k].conv0
                t1, rr = acc[row]
                d = saved.d1[k] if which == 1 else saved.d0[k]
                sp, ss = self._srow(styles, row)
                _lib.call('smc_sgrad_finish', _lib.ptr(t1), _lib.ptr(rr), _lib.ptr(L.q), _lib.ptr(d), sp, ss, _lib.ptr(gscale),
                          _lib.ptr(grad[i]), n, L.cin, L.cout, _lib.ptr(grad_samples[0, i]) if per_sample else None,
                          grad_samples.stride(0) if per_sample else 0, _lib.stream())
        return (grad, grad_samples) if per_sample else grad
