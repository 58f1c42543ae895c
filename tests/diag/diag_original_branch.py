"""Dev diagnostic (GPU box): can the no-gradient original-image branch (find_direction.py:312) run in a cheaper engine mode than the
edited branch?  Loss and delta-S gradient against the reference goldens (config1 = 256 px, config4 = 1024 px) with the edited branch at
x3p and the original branch at x3p / x3 / x1, at the golden's delta (|delta| ~ 0.1 per element) and at a 20x smaller delta -- the
directional loss normalises e = f(edited) - f(original), so an error in f(original) that does not cancel against the same error in
f(edited) is divided by |e|, which shrinks with delta."""
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
for name, res in (('config1', 256), ('config4', 1024)):
    g = np.load(f'tests/golden/{name}.npz')
    G = o_syn.make_generator(res, seed=0)
    ws = torch.as_tensor(g['ws'])
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    for scale in (1.0, 0.05):
        delta = torch.as_tensor(g['delta']) * scale
        if scale == 1.0:
            ref_loss, ref_grad = float(g['loss']), torch.as_tensor(g['grad'])[0]
        elif res == 256:          # small delta: the oracle on the CPU (256 px only; seconds)
            r = o_dir.direction_step(G, shapes, o_dir.CLIPLoss(o_vit.CLIP(o_vit.random_clip_params(seed=0)), pos, neg), S, delta, o_dir.RESOLUTION_TO_K[res])
            ref_loss, ref_grad = float(r['loss']), r['grad'][0]
        else:
            ref_loss = ref_grad = None
        base = None
        for op in ('x3p', 'x3', 'x1'):
            f = direction.DirectionFinder(G, model, pos, neg, res, precision='x3p', micro_batch=4, original_precision=op)
            f.delta.copy_(delta.cuda())
            out = f.step(S.cuda(), lr=0.0)
            loss, grad = out['loss'].item(), out['grad'].cpu()
            if base is None:
                base = (loss, grad)
            msg = f'{name} delta x{scale}: original branch {op}: loss {loss:.7f}'
            if ref_loss is not None:
                msg += f' rel err {abs(loss - ref_loss) / abs(ref_loss):.2e}; grad rel-l2 vs reference {((grad - ref_grad).norm() / ref_grad.norm()).item():.2e}'
            msg += f'; vs x3p original: loss {abs(loss - base[0]) / abs(base[0]):.2e} grad {((grad - base[1]).norm() / base[1].norm()).item():.2e}'
            print(msg, flush=True)
