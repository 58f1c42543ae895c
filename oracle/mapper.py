"""Oracle: the latent mapper and its training objective  (TEST INFRASTRUCTURE ONLY).

Restates ``latent_mappers.py:12-93`` (ModulationModule with ``embedding=None``, SubMapperModulation, Mapper; ``PixelNorm`` of
``encoder4editing/models/stylegan2/model.py:10-15``) as a function of a flat parameter dict under the reference's ``state_dict`` keys, and the
loss of one ``train_latent_mapper.py:150-176`` step (``find_direction.compute_loss``: CLIP term + identity term + L2 on the per-image delta;
landmarks coefficient 0).  ``pin_reference.py`` loads the dict into the REAL ``Mapper`` and runs the reference's real ``compute_loss``.
"""
import math

import torch
import torch.nn.functional as F

from . import direction, idloss, synthesis


def random_mapper_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / math.sqrt(512)
    p = {}
    for half in ('course_mapping', 'medium_mapping'):
        for i in range(5):
            k = f'{half}.modulation_module_list.{i}.fc.'
            p[k + 'weight'] = (torch.rand(512, 512, generator=g) * 2 - 1) * bound
            p[k + 'bias'] = (torch.rand(512, generator=g) * 2 - 1) * bound
    return p


def mapper_forward(p, x, neg_slope=0.01):
    """Mapper.forward (latent_mappers.py:75-93): x [B, 8, 512] -> delta [B, 8, 512]."""
    outs = []
    for half, xs in (('course_mapping', x[:, :4]), ('medium_mapping', x[:, 4:8])):
        h = xs * torch.rsqrt(torch.mean(xs ** 2, dim=1, keepdim=True) + 1e-8)            # PixelNorm over dim 1 (the four S rows)
        for i in range(5):                                                              # ModulationModule.forward, embedding=None (:21-31)
            k = f'{half}.modulation_module_list.{i}.fc.'
            h = F.linear(h, p[k + 'weight'], p[k + 'bias'])
            h = F.layer_norm(h, [4, 512])
            h = F.leaky_relu(h, neg_slope)
        outs.append(h)
    return torch.cat(outs, dim=1)


def mapper_step_loss(G, temp_shapes, clip_loss, p, styles, until_k, id_params=None, identity_loss_coef=0.0, clip_loss_coef=1.0, l2_reg_coef=0.1,
                     neg_slope=0.01, nada_prompts=None):
    """Loss of train_latent_mapper.py:150-176 as a differentiable function of the mapper parameters ``p`` (landmarks_loss_coef = 0)."""
    rows = direction.S_TRAINABLE_ROWS
    delta = mapper_forward(p, styles[:, rows], neg_slope)
    styles2 = styles.clone()
    styles2[:, rows] = styles2[:, rows] + delta
    _, img = synthesis.generate_image(G, until_k, styles2, temp_shapes)
    with torch.no_grad():
        _, original = synthesis.generate_image(G, until_k, styles, temp_shapes)
    if nada_prompts is not None:
        clip_term = clip_loss_coef * clip_loss(original, nada_prompts[0], img, nada_prompts[1])
    else:
        clip_term = clip_loss_coef * clip_loss(direction.unprocess(original), direction.unprocess(img))
    id_term = identity_loss_coef * idloss.id_loss(id_params, img, original) if identity_loss_coef else torch.zeros([])
    l2 = l2_reg_coef * F.mse_loss(styles2[:, rows], styles[:, rows])
    return dict(loss=clip_term + id_term + l2, clip_loss=clip_term, identity_loss=id_term, l2_loss=l2, delta=delta)
