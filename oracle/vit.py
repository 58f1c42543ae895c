"""Oracle: CLIP ViT-B/32 (and ViT-B/16) image tower + text tower  (TEST INFRASTRUCTURE ONLY).

Restates openai/CLIP ``clip/model.py`` (VisionTransformer, Transformer, ResidualAttentionBlock,
fp32-internal LayerNorm, QuickGELU = x*sigmoid(1.702x), CLIP.encode_image / encode_text).  The
package is NOT vendored in the reference (unpinned ``pip install git+https://github.com/openai/CLIP``,
reference ``README.md:13``) -- PARITY UNPINNED against upstream source; the call sites it must
serve are ``clip_loss.py:11-16,25-26``.  ``pin_reference.py`` cross-checks this restatement by
weight copy against ``transformers.CLIPModel`` (same architecture, independent code).

Parameter names follow the openai state_dict (``visual.conv1.weight``,
``visual.transformer.resblocks.{i}.attn.in_proj_weight`` ...), stored flat in ``self.p``.
"""
import math

import torch
import torch.nn.functional as F

VIT_B32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32,
               context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
               transformer_layers=12)
# "ViT-B/16" (clip_loss.py:12-13, the second tower of clip_type='double'): 16-px patches -> 14 x 14 + 1 = 197 tokens; the
# rest of the architecture is identical
VIT_B16 = dict(VIT_B32, vision_patch_size=16)


def _block_params(g, prefix, width, layers, p):
    # std scheme of CLIP.initialize_parameters (applied to BOTH towers here; upstream leaves the
    # visual blocks at nn defaults -- irrelevant for random-init parity work).
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    attn_std = width ** -0.5
    fc_std = (2 * width) ** -0.5
    for i in range(layers):
        b = f'{prefix}.resblocks.{i}.'
        p[b + 'ln_1.weight'] = 1 + 0.1 * torch.randn(width, generator=g)
        p[b + 'ln_1.bias'] = 0.1 * torch.randn(width, generator=g)
        p[b + 'attn.in_proj_weight'] = attn_std * torch.randn(3 * width, width, generator=g)
        p[b + 'attn.in_proj_bias'] = 0.02 * torch.randn(3 * width, generator=g)
        p[b + 'attn.out_proj.weight'] = proj_std * torch.randn(width, width, generator=g)
        p[b + 'attn.out_proj.bias'] = 0.02 * torch.randn(width, generator=g)
        p[b + 'ln_2.weight'] = 1 + 0.1 * torch.randn(width, generator=g)
        p[b + 'ln_2.bias'] = 0.1 * torch.randn(width, generator=g)
        p[b + 'mlp.c_fc.weight'] = fc_std * torch.randn(4 * width, width, generator=g)
        p[b + 'mlp.c_fc.bias'] = 0.02 * torch.randn(4 * width, generator=g)
        p[b + 'mlp.c_proj.weight'] = proj_std * torch.randn(width, 4 * width, generator=g)
        p[b + 'mlp.c_proj.bias'] = 0.02 * torch.randn(width, generator=g)


def random_clip_params(seed=0, cfg=VIT_B32):
    """Deterministic random-init state dict with the openai/CLIP key names (fp32)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    vw, ps = cfg['vision_width'], cfg['vision_patch_size']
    grid = cfg['image_resolution'] // ps
    sc = vw ** -0.5
    p['visual.conv1.weight'] = (3 * ps * ps) ** -0.5 * torch.randn(vw, 3, ps, ps, generator=g)
    p['visual.class_embedding'] = sc * torch.randn(vw, generator=g)
    p['visual.positional_embedding'] = sc * torch.randn(grid * grid + 1, vw, generator=g)
    p['visual.ln_pre.weight'] = 1 + 0.1 * torch.randn(vw, generator=g)
    p['visual.ln_pre.bias'] = 0.1 * torch.randn(vw, generator=g)
    _block_params(g, 'visual.transformer', vw, cfg['vision_layers'], p)
    p['visual.ln_post.weight'] = 1 + 0.1 * torch.randn(vw, generator=g)
    p['visual.ln_post.bias'] = 0.1 * torch.randn(vw, generator=g)
    p['visual.proj'] = sc * torch.randn(vw, cfg['embed_dim'], generator=g)
    tw = cfg['transformer_width']
    p['token_embedding.weight'] = 0.02 * torch.randn(cfg['vocab_size'], tw, generator=g)
    p['positional_embedding'] = 0.01 * torch.randn(cfg['context_length'], tw, generator=g)
    _block_params(g, 'transformer', tw, cfg['transformer_layers'], p)
    p['ln_final.weight'] = 1 + 0.1 * torch.randn(tw, generator=g)
    p['ln_final.bias'] = 0.1 * torch.randn(tw, generator=g)
    p['text_projection'] = tw ** -0.5 * torch.randn(tw, cfg['embed_dim'], generator=g)
    p['logit_scale'] = torch.tensor(math.log(1 / 0.07))       # clip/model.py CLIP.__init__ (no random draw: the other parameters are unchanged)
    return p


def synthetic_tokens(which, context_length=77):
    """Stand-in for ``clip.tokenize`` (no tokenizer offline): SOT 49406, some ids, EOT 49407 (unique
    max id, so ``text.argmax(-1)`` picks the EOT position as in upstream), zero padding."""
    body = {'pos': [320, 1125, 539, 320, 1710, 539, 320, 14387, 2308],
            'neg': [320, 1125, 539, 320, 1710, 539, 320, 25173, 786, 1237]}[which]
    t = torch.zeros(1, context_length, dtype=torch.int64)
    ids = [49406] + body + [49407]
    t[0, :len(ids)] = torch.tensor(ids)
    return t


def synthetic_tokenize(texts, context_length=77):
    """Stand-in for ``clip.tokenize(list of strings)`` (clip_loss_nada.py:117,131,224): one deterministic row per string, SOT, one id in
    [1000, 41000) per whitespace-separated word (crc32 of the word), EOT (the unique maximum, so ``argmax`` finds it), zero padding."""
    import zlib
    out = torch.zeros(len(texts), context_length, dtype=torch.int64)
    for r, text in enumerate(texts):
        ids = [49406] + [1000 + zlib.crc32(w.encode()) % 40000 for w in text.split()][:context_length - 2] + [49407]
        out[r, :len(ids)] = torch.tensor(ids)
    return out


def _ln(x, w, b):
    # clip/model.py LayerNorm: compute in fp32, cast back
    return F.layer_norm(x.float() if x.dtype == torch.float16 else x, (x.shape[-1],), w, b, 1e-5).to(x.dtype)


def _resblock(x, p, pre, heads, mask):
    """x [B,T,W].  x += MHA(ln_1(x)); x += c_proj(QuickGELU(c_fc(ln_2(x))))."""
    B, T, W = x.shape
    hd = W // heads
    h = _ln(x, p[pre + 'ln_1.weight'], p[pre + 'ln_1.bias'])
    qkv = h @ p[pre + 'attn.in_proj_weight'].t() + p[pre + 'attn.in_proj_bias']
    q, k, v = [t.reshape(B, T, heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
    att = (q * hd ** -0.5) @ k.transpose(-1, -2)
    if mask is not None:
        att = att + mask
    o = (att.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, T, W)
    x = x + (o @ p[pre + 'attn.out_proj.weight'].t() + p[pre + 'attn.out_proj.bias'])
    h = _ln(x, p[pre + 'ln_2.weight'], p[pre + 'ln_2.bias'])
    h = h @ p[pre + 'mlp.c_fc.weight'].t() + p[pre + 'mlp.c_fc.bias']
    h = h * torch.sigmoid(1.702 * h)
    return x + (h @ p[pre + 'mlp.c_proj.weight'].t() + p[pre + 'mlp.c_proj.bias'])


class CLIP:
    """Object with the surface ``clip_loss.py`` uses: ``encode_image``, ``encode_text``, ``dtype``."""

    def __init__(self, params=None, cfg=VIT_B32, dtype=torch.float32, seed=0):
        self.cfg = cfg
        src = params if params is not None else random_clip_params(seed, cfg)
        self.p = {k: v.to(dtype) for k, v in src.items()}
        self.dtype = dtype

    def encode_image(self, image):
        c, p = self.cfg, self.p
        x = F.conv2d(image.to(self.dtype), p['visual.conv1.weight'], stride=c['vision_patch_size'])
        x = x.flatten(2).transpose(1, 2)                                           # [B, 49, W]
        cls = p['visual.class_embedding'].expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + p['visual.positional_embedding']
        x = _ln(x, p['visual.ln_pre.weight'], p['visual.ln_pre.bias'])
        for i in range(c['vision_layers']):
            x = _resblock(x, p, f'visual.transformer.resblocks.{i}.', c['vision_width'] // 64, None)
        x = _ln(x[:, 0, :], p['visual.ln_post.weight'], p['visual.ln_post.bias'])
        return x @ p['visual.proj']

    def encode_text(self, text):
        c, p = self.cfg, self.p
        T = text.shape[1]
        x = p['token_embedding.weight'][text] + p['positional_embedding'][:T]
        mask = torch.full((T, T), float('-inf'), dtype=self.dtype, device=x.device).triu_(1)
        for i in range(c['transformer_layers']):
            x = _resblock(x, p, f'transformer.resblocks.{i}.', c['transformer_heads'], mask)
        x = _ln(x, p['ln_final.weight'], p['ln_final.bias'])
        return x[torch.arange(x.shape[0], device=x.device), text.argmax(dim=-1)] @ p['text_projection']


    def __call__(self, image, text):
        """clip/model.py CLIP.forward: (logits_per_image, logits_per_text) = exp(logit_scale) * cosine similarities."""
        i, t = self.encode_image(image), self.encode_text(text)
        i, t = i / i.norm(dim=1, keepdim=True), t / t.norm(dim=1, keepdim=True)
        logits = self.p['logit_scale'].exp() * i @ t.t()
        return logits, logits.t()


FLOPS_PER_IMAGE_FWD = None  # filled by bench from the GEMM shapes; see SURVEY.md section 8d
