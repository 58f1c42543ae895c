"""Dev diagnostic (GPU box): does the backward pass need lo planes for the activation GRADIENTS?  One find_direction step on the reference
goldens (step64, config1 = 256 px, config4 = 1024 px) with the gradient operands of the dgrad GEMMs as hi + lo planes (3 MMAs per product,
STYLEMC_GRAD_LO=1) and as a hi plane only (2 MMAs: A_hi*B_hi + A_hi*B_lo), at the golden's delta and at a 20x smaller one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle import direction as o_dir
from oracle import synthesis as o_syn
from oracle import vit as o_vit
from stylemc_b200 import clip, direction

model = clip.CLIPModel(o_vit.random_clip_params(seed=0), 'cuda', precision='x3p')
pos, neg = o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg')
for name, res, kw in (('step64', 64, dict(seed=1, channel_base=2048, channel_max=512)), ('config1', 256, dict(seed=0)), ('config4', 1024, dict(seed=0))):
    g = np.load(f'tests/golden/{name}.npz')
    G = o_syn.make_generator(res, **kw)
    if 'ws' in g:
        ws = torch.as_tensor(g['ws'])
        S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    else:
        S, shapes = torch.as_tensor(g['styles']), o_syn.get_temp_shapes(G)
    for scale in (1.0, 0.05):
        delta = torch.as_tensor(g['delta']) * scale
        ref_grad = torch.as_tensor(g['grad'])[0] if scale == 1.0 else None
        if scale != 1.0 and res <= 256:
            r = o_dir.direction_step(G, shapes, o_dir.CLIPLoss(o_vit.CLIP(o_vit.random_clip_params(seed=0)), pos, neg), S, delta, o_dir.RESOLUTION_TO_K.get(res, 100))
            ref_grad = r['grad'][0]
        base = None
        for lo, clo, bprec in ((True, True, 'x2'), (False, True, 'x2'), (False, False, 'x2'), (False, False, 'x1')):
            f = direction.DirectionFinder(G, model, pos, neg, res, precision='x3p', micro_batch=4)
            f.engine.grad_lo, model.grad_lo = lo, clo
            f.engine.bwd_prec = model.bwd_prec = bprec
            f.delta.copy_(delta.cuda())
            grad = f.step(S.cuda(), lr=0.0)['grad'].cpu()
            base = grad if base is None else base
            msg = f'{name} delta x{scale}: gradient planes synthesis {"hi+lo" if lo else "hi only"} / ViT {"hi+lo" if clo else "hi only"}, weights {"hi+lo" if bprec == "x2" else "hi only"}:'
            if ref_grad is not None:
                msg += f' grad rel-l2 vs reference {((grad - ref_grad).norm() / ref_grad.norm()).item():.3e};'
            print(msg + f' vs hi+lo {((grad - base).norm() / base.norm()).item():.3e}', flush=True)
