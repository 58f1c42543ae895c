"""Oracle: one find_direction optimisation step  (TEST INFRASTRUCTURE ONLY).

Restates, for clip_type='small' or 'double' / clip_loss_type='default' / identity = landmarks = 0:
``find_direction.py:38-41`` (trainable S rows), ``:49-52`` (unprocess), ``:148-169``
(compute_clip_loss), ``:190-191`` (L2 term), ``:298-339`` (cosine LR, delta insertion, two
synthesis passes, backward, SGD) and ``clip_loss.py:8-34`` (directional CLIP loss).

``Resize(224, interpolation=BICUBIC)`` of ``find_direction.py:258`` is, on tensors under
torchvision 0.26 (this image), ``F.interpolate(mode='bicubic', antialias=True,
align_corners=False)`` -- checked against the real ``Compose([Resize, CenterCrop])`` in
``pin_reference.py``.
"""
import math

import torch
import torch.nn.functional as F

from . import synthesis

S_TRAINABLE_ROWS = [2, 3, 5, 6, 8, 9, 11, 12]           # find_direction.py:41
RESOLUTION_TO_K = {256: 6, 512: 7, 1024: 8}              # find_direction.py:263
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)          # utils.py:91
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)          # utils.py:92


def unprocess(img, size=224):
    """find_direction.py:49-52: [-1,1]-ish image -> CLIP-normalised 224x224."""
    x = (img * 127.5 + 128).clamp(0, 255)
    x = F.interpolate(x, size=(size, size), mode='bicubic', antialias=True, align_corners=False)
    mean = torch.tensor(CLIP_MEAN, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    return (x / 255 - mean) / std


class CLIPLoss:
    """clip_loss.py:8-34 with an injected model (``clip.load`` is unavailable offline)."""

    def __init__(self, model, pos_tokens, neg_tokens):
        self.model = model
        t = model.encode_text(pos_tokens) - model.encode_text(neg_tokens)      # :15-17
        self.text_features = t / t.norm(dim=1, keepdim=True)                     # :18

    def __call__(self, src_image, tgt_image):
        e = self.model.encode_image(tgt_image) - self.model.encode_image(src_image)   # :25-27
        e = e / e.norm(dim=1, keepdim=True)                                            # :28
        cos = F.cosine_similarity(e, self.text_features)                               # :29-32
        return (len(src_image) - cos.sum()) / len(src_image)                           # :34


class DoubleCLIPLoss:
    """clip_type='double' (the CLI default, find_direction.py:216): ``init_clip_loss`` builds one CLIPLoss on ViT-B/32 and one
    on ViT-B/16 (:117-119), ``compute_clip_loss`` adds them as ``loss32 + 0.5 * loss16`` (:160-164)."""

    def __init__(self, loss_small, loss_large):
        self.loss_small, self.loss_large = loss_small, loss_large

    def __call__(self, src_image, tgt_image):
        return self.loss_small(src_image, tgt_image) + 0.5 * self.loss_large(src_image, tgt_image)


def cosine_lr(base_lr, it, total):
    """find_direction.py:298-299 (``it`` is 1-based)."""
    return math.cos(math.pi * it / total) * base_lr * 0.5 + base_lr * 0.5


def direction_step(G, temp_shapes, clip_loss, styles, delta, until_k, clip_loss_coef=1.0, l2_reg_coef=0.1,
                   noise_mode='const'):
    """One loss/gradient evaluation of find_direction.py:306-336.

    styles [N,26,512]; delta [1,8,512] (leaf).  Returns dict(loss, clip_loss, l2_loss, grad, img,
    original_img).  grad is d loss / d delta, [1,8,512].
    """
    delta = delta.detach().clone().requires_grad_(True)
    direction = torch.zeros(1, synthesis.N_STYLE_ROWS, synthesis.STYLE_WIDTH, dtype=styles.dtype, device=styles.device)
    direction = direction.index_put((torch.tensor([0], device=styles.device).view(1, 1), torch.tensor(S_TRAINABLE_ROWS, device=styles.device).view(1, -1)),
                                    delta)                                          # :307
    styles2 = styles + direction                                                   # :308
    _, img = synthesis.generate_image(G, until_k, styles2, temp_shapes, noise_mode)   # :309
    with torch.no_grad():
        _, original = synthesis.generate_image(G, until_k, styles, temp_shapes, noise_mode)  # :312
    clip_term = clip_loss_coef * clip_loss(unprocess(original), unprocess(img))   # :159-169
    l2_term = l2_reg_coef * F.mse_loss(styles2[:, S_TRAINABLE_ROWS], styles[:, S_TRAINABLE_ROWS])  # :190-191
    loss = clip_term + l2_term
    grad, = torch.autograd.grad(loss, delta)
    return dict(loss=loss.detach(), clip_loss=clip_term.detach(), l2_loss=l2_term.detach(), grad=grad,
                img=img.detach(), original_img=original)


def sgd_update(delta, grad, lr):
    """torch.optim.SGD without momentum, find_direction.py:285,339."""
    return delta - lr * grad
