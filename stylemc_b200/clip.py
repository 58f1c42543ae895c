"""CLIP ViT-B/32 (and ViT-B/16) image tower (forward + backward to the pixels) and text tower (forward) on hand-written kernels.

Stands in for the object ``clip.load(...)`` returns at clip_loss.py:11: ``encode_image(x[N,3,224,224]) -> [N,512]``,
``encode_text(tokens[B,77]) -> [B,512]``, ``dtype`` (openai/CLIP clip/model.py VisionTransformer / Transformer /
ResidualAttentionBlock / QuickGELU; parameter names are the openai state-dict keys).  Every linear layer is one launch of the
tcgen05 implicit GEMM (``smc_igemm``) with bias and residual add in its epilogue; LayerNorm, the attention core (50 tokens for
ViT-B/32 and 77 for the text tower in one CTA per head; 197 tokens for ViT-B/16 in row blocks) and QuickGELU are small fused kernels that emit the next GEMM's fp16 operand planes directly.  The residual stream and all
statistics stay fp32.  Weights are frozen: the backward pass produces the gradient w.r.t. the input pixels only (the
reference accumulates 88 M unused weight gradients per step).
"""
import os

import torch

from . import _lib, gemm

VIT_B32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32,
               context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8, transformer_layers=12)
# clip.load("ViT-B/16") (clip_loss.py:12-13), the second tower of clip_type='double': 16-px patches, 14 * 14 + 1 = 197 tokens
VIT_B16 = dict(VIT_B32, vision_patch_size=16)


def whole_sequence_attention_bwd_fits(t, head_dim):
    """smc_attention_bwd keeps q, k, v, dO and two [t, t+1] score matrices of one head in shared memory (200 KB budget, vit.cu); longer
    sequences take smc_attention_bwd_tiled."""
    return (4 * t * (head_dim + 1) + 2 * t * (t + 1)) * 4 <= 200 * 1024


class _Linear:
    """Frozen nn.Linear weights as pre-scaled fp16 planes for y = x @ W.T (fwd) and dx = dy @ W (bwd)."""

    def __init__(self, w, b, device, two, need_bwd):
        w = w.detach().to(device=device, dtype=torch.float32)
        self.out_f, self.in_f = w.shape
        # [P*out, in] (y = x @ W.T) and [P*in, out] (dx = dy @ W) planes, pre-scaled by a power of two, in one kernel
        self.fwd, self.bwd, _, self.inv = gemm.prepare_weights(w, two=two, fwd=True, bwd=need_bwd, prescale=True)
        self.bias = b.detach().to(device=device, dtype=torch.float32).contiguous() if b is not None else None


class _Tower:
    def __init__(self, p, prefix, width, layers, heads, device, two, need_bwd):
        self.width, self.layers, self.heads = width, layers, heads
        f = lambda k: p[k].detach().to(device=device, dtype=torch.float32).contiguous()
        self.blocks = []
        for i in range(layers):
            b = f'{prefix}.resblocks.{i}.'
            self.blocks.append(dict(
                ln1=(f(b + 'ln_1.weight'), f(b + 'ln_1.bias')), ln2=(f(b + 'ln_2.weight'), f(b + 'ln_2.bias')),
                qkv=_Linear(p[b + 'attn.in_proj_weight'], p[b + 'attn.in_proj_bias'], device, two, need_bwd),
                out=_Linear(p[b + 'attn.out_proj.weight'], p[b + 'attn.out_proj.bias'], device, two, need_bwd),
                fc=_Linear(p[b + 'mlp.c_fc.weight'], p[b + 'mlp.c_fc.bias'], device, two, need_bwd),
                proj=_Linear(p[b + 'mlp.c_proj.weight'], p[b + 'mlp.c_proj.bias'], device, two, need_bwd)))


class CLIPModel:
    """precision: 'x3p' (split-fp16 operands + promoted accumulation: fp32 accuracy; default), 'x3' (split operands) or
    'x1' (fp16 operands)."""

    def __init__(self, params, device='cuda', precision='x3p', cfg=VIT_B32):
        if precision not in ('x1', 'x3', 'x3p'):
            raise ValueError(precision)
        self.acc_k = 512 if precision == 'x3p' else 0
        precision = 'x3' if precision == 'x3p' else precision
        self.cfg, self.device, self.precision = cfg, torch.device(device), precision
        self.two = precision == 'x3'
        # backward GEMMs: the gradient operand keeps its lo plane here (3 MMAs per product).  The synthesis engine drops it (2 MMAs, +8 %
        # images/s); in the ViT the same switch (STYLEMC_CLIP_GRAD_LO=0) gains nothing measurable (the tower is 5 % of the step) and costs
        # 1.5e-4 of gradient accuracy on the 1024-px golden, so it stays on
        self.grad_lo = os.environ.get('STYLEMC_CLIP_GRAD_LO', '1') != '0'
        self.bwd_prec = os.environ.get('STYLEMC_BWD_PREC', 'x2')
        self.dtype = torch.float32
        dev, two = self.device, self.two
        f = lambda k: params[k].detach().to(device=dev, dtype=torch.float32).contiguous()
        vw, ps = cfg['vision_width'], cfg['vision_patch_size']
        self.conv1 = _Linear(params['visual.conv1.weight'].reshape(vw, 3 * ps * ps), None, dev, two, True)
        self.cls, self.pos = f('visual.class_embedding'), f('visual.positional_embedding')
        self.ln_pre = (f('visual.ln_pre.weight'), f('visual.ln_pre.bias'))
        self.ln_post = (f('visual.ln_post.weight'), f('visual.ln_post.bias'))
        self.vproj = f('visual.proj')
        self.visual = _Tower(params, 'visual.transformer', vw, cfg['vision_layers'], vw // 64, dev, two, True)
        self.tok_emb, self.tpos = f('token_embedding.weight'), f('positional_embedding')
        self.ln_final = (f('ln_final.weight'), f('ln_final.bias'))
        self.tproj = f('text_projection')
        self.logit_scale = float(params['logit_scale']) if 'logit_scale' in params else 4.605170185988092 - 1.9459101090932196   # ln(1 / 0.07)
        self.text = _Tower(params, 'transformer', cfg['transformer_width'], cfg['transformer_layers'], cfg['transformer_heads'], dev, two, False)

    # ---- kernels ---------------------------------------------------------------------------------
    def _planes(self, rows, width):
        return torch.empty([2 if self.two else 1, rows, width], dtype=torch.float16, device=self.device)

    def _lo(self, planes):
        return _lib.ptr(planes[1]) if (self.two and planes.shape[0] == 2) else None

    def _gplanes(self, rows, width):
        """Operand planes of a backward GEMM (gradient): hi (+ lo only with ``grad_lo``)."""
        return torch.empty([2 if (self.two and self.grad_lo) else 1, rows, width], dtype=torch.float16, device=self.device)

    def _linear(self, a_planes, lin, rows, bwd=False, residual=None, out=None):
        """out[rows, n_out] = A @ B.T (+ bias) (+ residual), fp32."""
        B = lin.bwd if bwd else lin.fwd
        n_out, k = (lin.in_f, lin.out_f) if bwd else (lin.out_f, lin.in_f)
        if out is None:
            out = torch.empty([rows, n_out], dtype=torch.float32, device=self.device)
        precision = self.bwd_prec if (self.two and a_planes.shape[0] == 1) else self.precision      # hi-only A operand: two-term split
        gemm.igemm(a_planes.reshape(-1, 1, rows, k), B, 1, 1, rows, n_out, gemm.TAPS_1X1, precision=precision,
                   a_plane_stride_imgs=1, b_rows_per_tap=n_out, bias=None if bwd else lin.bias, residual=residual, out_f32=out,
                   acc_scale=lin.inv, acc_chunk_k=self.acc_k)
        return out

    def _ln(self, x, wb, rows, width, stride=1, offset=0, want32=False, stats=True):
        mean = torch.empty(rows, dtype=torch.float32, device=self.device) if stats else None
        rstd = torch.empty(rows, dtype=torch.float32, device=self.device) if stats else None
        y32 = torch.empty([rows, width], dtype=torch.float32, device=self.device) if want32 else None
        planes = None if want32 else self._planes(rows, width)
        _lib.call('smc_layernorm_fwd', _lib.ptr(x), stride, offset, _lib.ptr(wb[0]), _lib.ptr(wb[1]), _lib.ptr(y32),
                  None if want32 else _lib.ptr(planes[0]), None if want32 else self._lo(planes), _lib.ptr(mean), _lib.ptr(rstd), rows, width,
                  _lib.stream())
        return (y32 if want32 else planes), mean, rstd

    def _tower_fwd(self, tower, x, b, t, causal, save):
        rows, wd = b * t, tower.width
        saved = []
        for blk in tower.blocks:
            h, m1, r1 = self._ln(x, blk['ln1'], rows, wd)
            qkv = self._linear(h, blk['qkv'], rows)
            o = self._planes(rows, wd)
            _lib.call('smc_attention_fwd', _lib.ptr(qkv), _lib.ptr(o[0]), self._lo(o), None, b, t, wd, tower.heads, int(causal), _lib.stream())
            x_mid = self._linear(o, blk['out'], rows, residual=x)
            h2, m2, r2 = self._ln(x_mid, blk['ln2'], rows, wd)
            hfc = self._linear(h2, blk['fc'], rows)
            g = self._planes(rows, 4 * wd)
            _lib.call('smc_quickgelu_fwd', _lib.ptr(hfc), _lib.ptr(g[0]), self._lo(g), rows * 4 * wd, _lib.stream())
            x_out = self._linear(g, blk['proj'], rows, residual=x_mid)
            if save:
                saved.append(dict(x_in=x, m1=m1, r1=r1, qkv=qkv, x_mid=x_mid, m2=m2, r2=r2, hfc=hfc))
            x = x_out
        return x, saved

    # ---- image tower -----------------------------------------------------------------------------
    def encode_image_fwd(self, image, save=False):
        """image [B, 3, 224, 224] fp32 CUDA -> (features [B, 512], saved state for encode_image_bwd or None)."""
        _lib.require_cuda(image, 'image')
        cfg = self.cfg
        res, ps, wd = cfg['image_resolution'], cfg['vision_patch_size'], cfg['vision_width']
        if image.ndim != 4 or tuple(image.shape[1:]) != (3, res, res):
            raise RuntimeError(f'encode_image expects [B, 3, {res}, {res}], got {tuple(image.shape)}')
        image = image.float().contiguous()
        b = image.shape[0]
        grid = res // ps
        t = grid * grid + 1
        with torch.cuda.device(self.device):
            cols = self._planes(b * grid * grid, 3 * ps * ps)
            _lib.call('smc_patchify', _lib.ptr(image), _lib.ptr(cols[0]), self._lo(cols), b, res, ps, _lib.stream())
            patch = self._linear(cols, self.conv1, b * grid * grid)
            x0 = torch.empty([b * t, wd], dtype=torch.float32, device=self.device)
            _lib.call('smc_assemble_tokens', _lib.ptr(patch), _lib.ptr(self.cls), _lib.ptr(self.pos), _lib.ptr(x0), b, t, wd, _lib.stream())
            x, m0, r0 = self._ln(x0, self.ln_pre, b * t, wd, want32=True)
            xl, saved_blocks = self._tower_fwd(self.visual, x, b, t, False, save)
            ln, mp, rp = self._ln(xl, self.ln_post, b, wd, stride=t, offset=0, want32=True)
            feat = torch.empty([b, cfg['embed_dim']], dtype=torch.float32, device=self.device)
            _lib.call('smc_head_proj', _lib.ptr(ln), _lib.ptr(self.vproj), _lib.ptr(feat), b, wd, cfg['embed_dim'], _lib.stream())
        saved = dict(b=b, t=t, x0=x0, m0=m0, r0=r0, blocks=saved_blocks, xl=xl, mp=mp, rp=rp) if save else None
        return feat, saved

    def encode_image_bwd(self, saved, d_feat):
        """d_feat [B, 512] (may carry a loss scale) -> gradient w.r.t. the input image [B, 3, 224, 224] (same scale)."""
        cfg = self.cfg
        res, ps, wd = cfg['image_resolution'], cfg['vision_patch_size'], cfg['vision_width']
        b, t = saved['b'], saved['t']
        rows = b * t
        grid = res // ps
        tw = self.visual
        with torch.cuda.device(self.device):
            d_feat = d_feat.float().contiguous()
            dln = torch.empty([b, wd], dtype=torch.float32, device=self.device)
            _lib.call('smc_head_proj_bwd', _lib.ptr(d_feat), _lib.ptr(self.vproj), _lib.ptr(dln), b, wd, cfg['embed_dim'], _lib.stream())
            tiled = not whole_sequence_attention_bwd_fits(t, wd // tw.heads) or bool(int(os.environ.get('STYLEMC_ATTENTION_TILED', '0')))
            stats = torch.empty([2, b * tw.heads * t], dtype=torch.float32, device=self.device) if tiled else None
            dx = torch.zeros([rows, wd], dtype=torch.float32, device=self.device)
            _lib.call('smc_layernorm_bwd', _lib.ptr(dln), _lib.ptr(saved['xl']), t, 0, _lib.ptr(self.ln_post[0]), _lib.ptr(saved['mp']),
                      _lib.ptr(saved['rp']), _lib.ptr(dx), b, wd, 0, _lib.stream())
            for blk, sv in zip(reversed(tw.blocks), reversed(saved['blocks'])):
                # MLP branch
                dxp = self._gplanes(rows, wd)
                _lib.call('smc_split_rows', _lib.ptr(dx), _lib.ptr(dxp[0]), self._lo(dxp), rows, wd, rows, 0, 0, _lib.stream())
                dg = self._linear(dxp, blk['proj'], rows, bwd=True)
                dh = self._gplanes(rows, 4 * wd)
                _lib.call('smc_quickgelu_bwd', _lib.ptr(dg), _lib.ptr(sv['hfc']), _lib.ptr(dh[0]), self._lo(dh), rows * 4 * wd, _lib.stream())
                dln2 = self._linear(dh, blk['fc'], rows, bwd=True)
                _lib.call('smc_layernorm_bwd', _lib.ptr(dln2), _lib.ptr(sv['x_mid']), 1, 0, _lib.ptr(blk['ln2'][0]), _lib.ptr(sv['m2']),
                          _lib.ptr(sv['r2']), _lib.ptr(dx), rows, wd, 1, _lib.stream())
                # attention branch
                _lib.call('smc_split_rows', _lib.ptr(dx), _lib.ptr(dxp[0]), self._lo(dxp), rows, wd, rows, 0, 0, _lib.stream())
                do = self._linear(dxp, blk['out'], rows, bwd=True)
                dqkv = self._gplanes(rows, 3 * wd)
                if tiled:
                    _lib.call('smc_attention_bwd_tiled', _lib.ptr(sv['qkv']), _lib.ptr(do), _lib.ptr(dqkv[0]), self._lo(dqkv), _lib.ptr(stats), b, t,
                              wd, tw.heads, 0, _lib.stream())
                else:
                    _lib.call('smc_attention_bwd', _lib.ptr(sv['qkv']), _lib.ptr(do), _lib.ptr(dqkv[0]), self._lo(dqkv), b, t, wd, tw.heads, 0,
                              _lib.stream())
                dln1 = self._linear(dqkv, blk['qkv'], rows, bwd=True)
                _lib.call('smc_layernorm_bwd', _lib.ptr(dln1), _lib.ptr(sv['x_in']), 1, 0, _lib.ptr(blk['ln1'][0]), _lib.ptr(sv['m1']),
                          _lib.ptr(sv['r1']), _lib.ptr(dx), rows, wd, 1, _lib.stream())
            dx0 = torch.empty([rows, wd], dtype=torch.float32, device=self.device)
            _lib.call('smc_layernorm_bwd', _lib.ptr(dx), _lib.ptr(saved['x0']), 1, 0, _lib.ptr(self.ln_pre[0]), _lib.ptr(saved['m0']),
                      _lib.ptr(saved['r0']), _lib.ptr(dx0), rows, wd, 0, _lib.stream())
            prow = b * grid * grid
            dpp = self._gplanes(prow, wd)
            _lib.call('smc_split_rows', _lib.ptr(dx0), _lib.ptr(dpp[0]), self._lo(dpp), prow, wd, t - 1, t, 1, _lib.stream())
            dcols = self._linear(dpp, self.conv1, prow, bwd=True)
            dimg = torch.empty([b, 3, res, res], dtype=torch.float32, device=self.device)
            _lib.call('smc_unpatchify', _lib.ptr(dcols), _lib.ptr(dimg), b, res, ps, _lib.stream())
        return dimg

    def encode_image(self, image):
        """clip_loss.py:25-26 entry point; differentiable w.r.t. ``image`` through torch autograd."""
        return _EncodeImage.apply(image, self)

    # ---- text tower ------------------------------------------------------------------------------
    def encode_text(self, text):
        """text [B, 77] int64 token ids -> [B, 512] (clip_loss.py:15-16).  Forward only (computed once per prompt)."""
        cfg = self.cfg
        text = text.to(self.device, torch.int64).contiguous()
        b, t = text.shape
        if t > cfg['context_length']:
            raise RuntimeError('text longer than the context length')
        wd = cfg['transformer_width']
        eot = text.argmax(dim=-1).tolist()
        with torch.cuda.device(self.device):
            x0 = torch.empty([b * t, wd], dtype=torch.float32, device=self.device)
            _lib.call('smc_embed_text', _lib.ptr(text), _lib.ptr(self.tok_emb), _lib.ptr(self.tpos), _lib.ptr(x0), b, t, wd, _lib.stream())
            x, _ = self._tower_fwd(self.text, x0, b, t, True, False)
            feat = torch.empty([b, cfg['embed_dim']], dtype=torch.float32, device=self.device)
            for i in range(b):
                ln, _, _ = self._ln(x, self.ln_final, 1, wd, stride=1, offset=i * t + eot[i], want32=True, stats=False)
                _lib.call('smc_head_proj', _lib.ptr(ln), _lib.ptr(self.tproj), _lib.ptr(feat[i]), 1, wd, cfg['embed_dim'], _lib.stream())
        return feat


class _EncodeImage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, model):
        feat, saved = model.encode_image_fwd(image, save=image.requires_grad)
        ctx.saved, ctx.model = saved, model
        return feat

    @staticmethod
    def backward(ctx, d_feat):
        if ctx.saved is None:
            return None, None
        # scale the (usually tiny) embedding gradient into fp16 range for the backward GEMMs, and back afterwards
        amax = float(d_feat.abs().max())
        k = 2.0 ** (-torch.tensor(amax).log2().floor().item()) if amax > 0 else 1.0
        return ctx.model.encode_image_bwd(ctx.saved, d_feat * k) / k, None


def random_params(seed=0, cfg=VIT_B32):
    """Random-init state dict with the openai/CLIP key names (BASELINE.json: random-init CLIP ViT-B/32; no checkpoints offline).
    Scales follow CLIP.initialize_parameters (attn width^-0.5, proj (2*layers*width)^-0.5, fc (2*width)^-0.5)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *shape: torch.randn(*shape, generator=g)
    p = {}
    vw, ps, tw = cfg['vision_width'], cfg['vision_patch_size'], cfg['transformer_width']
    grid = cfg['image_resolution'] // ps

    def tower(prefix, width, layers):
        attn, proj, fc = width ** -0.5, (2 * layers * width) ** -0.5, (2 * width) ** -0.5
        for i in range(layers):
            b = f'{prefix}.resblocks.{i}.'
            for ln in ('ln_1', 'ln_2'):
                p[b + ln + '.weight'], p[b + ln + '.bias'] = 1 + 0.1 * rn(width), 0.1 * rn(width)
            p[b + 'attn.in_proj_weight'], p[b + 'attn.in_proj_bias'] = attn * rn(3 * width, width), 0.02 * rn(3 * width)
            p[b + 'attn.out_proj.weight'], p[b + 'attn.out_proj.bias'] = proj * rn(width, width), 0.02 * rn(width)
            p[b + 'mlp.c_fc.weight'], p[b + 'mlp.c_fc.bias'] = fc * rn(4 * width, width), 0.02 * rn(4 * width)
            p[b + 'mlp.c_proj.weight'], p[b + 'mlp.c_proj.bias'] = proj * rn(width, 4 * width), 0.02 * rn(width)

    p['visual.conv1.weight'] = (3 * ps * ps) ** -0.5 * rn(vw, 3, ps, ps)
    p['visual.class_embedding'] = vw ** -0.5 * rn(vw)
    p['visual.positional_embedding'] = vw ** -0.5 * rn(grid * grid + 1, vw)
    p['visual.ln_pre.weight'], p['visual.ln_pre.bias'] = 1 + 0.1 * rn(vw), 0.1 * rn(vw)
    tower('visual.transformer', vw, cfg['vision_layers'])
    p['visual.ln_post.weight'], p['visual.ln_post.bias'] = 1 + 0.1 * rn(vw), 0.1 * rn(vw)
    p['visual.proj'] = vw ** -0.5 * rn(vw, cfg['embed_dim'])
    p['token_embedding.weight'] = 0.02 * rn(cfg['vocab_size'], tw)
    p['positional_embedding'] = 0.01 * rn(cfg['context_length'], tw)
    tower('transformer', tw, cfg['transformer_layers'])
    p['ln_final.weight'], p['ln_final.bias'] = 1 + 0.1 * rn(tw), 0.1 * rn(tw)
    p['text_projection'] = tw ** -0.5 * rn(tw, cfg['embed_dim'])
    p['logit_scale'] = torch.tensor(2.6592600369327783)        # ln(1 / 0.07), CLIP.__init__
    return p


def placeholder_tokens(body, context_length=77):
    """Token row for a prompt when ``clip.tokenize`` (BPE vocabulary) is unavailable: <SOT> body... <EOT> zero padding; EOT
    (49407) is the largest id, which is what ``encode_text`` keys on (text.argmax(-1))."""
    ids = [49406] + [int(t) for t in body] + [49407]
    if len(ids) > context_length:
        raise RuntimeError('prompt longer than the context length')
    t = torch.zeros(1, context_length, dtype=torch.int64)
    t[0, :len(ids)] = torch.tensor(ids)
    return t
