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


# clip_loss_nada.py:12-40 (the prompt-engineering templates every class string is composed with)
NADA_TEMPLATES = [
    'a photo of a {}.', 'a rendering of a {}.', 'a cropped photo of the {}.', 'the photo of a {}.', 'a photo of a clean {}.', 'a photo of a dirty {}.',
    'a dark photo of the {}.', 'a photo of my {}.', 'a photo of the cool {}.', 'a close-up photo of a {}.', 'a bright photo of the {}.',
    'a cropped photo of a {}.', 'a photo of the {}.', 'a good photo of the {}.', 'a photo of one {}.', 'a close-up photo of the {}.',
    'a rendition of the {}.', 'a photo of the clean {}.', 'a rendition of a {}.', 'a photo of a nice {}.', 'a good photo of a {}.',
    'a photo of the nice {}.', 'a photo of the small {}.', 'a photo of the weird {}.', 'a photo of the large {}.', 'a photo of a cool {}.',
    'a photo of a small {}.']


def nada_preprocess(img, size=224):
    """``CLIPLoss.preprocess`` of clip_loss_nada.py:86-89 on a square tensor: Normalize(mean -1, std 2) (GAN output -> [0, 1], NO clamp),
    Resize(224, BICUBIC) + CenterCrop(224) (``clip_preprocess.transforms[:2]``), Normalize(CLIP mean, std) (``transforms[4:]``)."""
    x = (img + 1.0) / 2.0
    x = F.interpolate(x, size=(size, size), mode='bicubic', antialias=True, align_corners=False)
    mean = torch.tensor(CLIP_MEAN, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    return (x - mean) / std


class CLIPLossNADA:
    """clip_loss_nada.py:66-345 for the two configurations find_direction.py:101-114 builds: ``lambda_direction=1`` (clip_loss_type 'nada':
    ``clip_directional_loss``, :206-218) or ``lambda_global=1, lambda_direction=0`` ('nada_global': ``global_clip_loss``, :220-229).  Called
    as the reference calls it (find_direction.py:151-158): ``loss(original, negative_prompt, generated, prompt)`` on raw GAN outputs."""

    def __init__(self, model, tokenize, lambda_direction=1.0, lambda_global=0.0):
        self.model, self.tokenize = model, tokenize
        self.lambda_direction, self.lambda_global = lambda_direction, lambda_global
        self.target_direction = None

    def get_text_features(self, class_str):                                     # :129-140
        f = self.model.encode_text(self.tokenize([t.format(class_str) for t in NADA_TEMPLATES])).detach()
        return f / f.norm(dim=-1, keepdim=True)

    def get_image_features(self, img):                                          # :142-148
        f = self.model.encode_image(nada_preprocess(img))
        return f / f.norm(dim=-1, keepdim=True)

    def compute_text_direction(self, source_class, target_class):               # :150-157
        d = (self.get_text_features(target_class) - self.get_text_features(source_class)).mean(dim=0, keepdim=True)
        return d / d.norm(dim=-1, keepdim=True)

    def clip_directional_loss(self, src_img, source_class, target_img, target_class):    # :206-218
        if self.target_direction is None:
            self.target_direction = self.compute_text_direction(source_class, target_class)
        e = self.get_image_features(target_img) - self.get_image_features(src_img)
        e = e / e.norm(dim=-1, keepdim=True)
        return (1.0 - F.cosine_similarity(e, self.target_direction)).mean()

    def global_clip_loss(self, img, text):                                      # :220-229
        logits_per_image, _ = self.model(nada_preprocess(img), self.tokenize(text))
        return (1.0 - logits_per_image / 100).mean()

    def __call__(self, src_img, source_class, target_img, target_class):        # :325-345
        loss = 0.0
        if self.lambda_global:
            loss = loss + self.lambda_global * self.global_clip_loss(target_img, [f'a {target_class}'])
        if self.lambda_direction:
            loss = loss + self.lambda_direction * self.clip_directional_loss(src_img, source_class, target_img, target_class)
        return loss


def cosine_lr(base_lr, it, total):
    """find_direction.py:298-299 (``it`` is 1-based)."""
    return math.cos(math.pi * it / total) * base_lr * 0.5 + base_lr * 0.5


def direction_step(G, temp_shapes, clip_loss, styles, delta, until_k, clip_loss_coef=1.0, l2_reg_coef=0.1,
                   noise_mode='const', nada_prompts=None):
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
    if nada_prompts is not None:       # find_direction.py:150-158: the NADA losses take the raw images and the (negative, positive) prompts
        clip_term = clip_loss_coef * clip_loss(original, nada_prompts[0], img, nada_prompts[1])
    else:
        clip_term = clip_loss_coef * clip_loss(unprocess(original), unprocess(img))   # :159-169
    l2_term = l2_reg_coef * F.mse_loss(styles2[:, S_TRAINABLE_ROWS], styles[:, S_TRAINABLE_ROWS])  # :190-191
    loss = clip_term + l2_term
    grad, = torch.autograd.grad(loss, delta)
    return dict(loss=loss.detach(), clip_loss=clip_term.detach(), l2_loss=l2_term.detach(), grad=grad,
                img=img.detach(), original_img=original)


def sgd_update(delta, grad, lr):
    """torch.optim.SGD without momentum, find_direction.py:285,339."""
    return delta - lr * grad
