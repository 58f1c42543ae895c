"""Gradient / loss error of one find_direction step against the golden fixtures for each CLIP precision mode (synthesis stays x3p).
usage (GPU): python tools/diag_clip_prec.py [x3p x3 x1]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synthesis as o_syn, vit as o_vit
from stylemc_b200 import clip, direction

def golden(name):
    return np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', name + '.npz'))

def run(tag, G, res, styles, g, cprec, **kw):
    model = clip.CLIPModel(o_vit.random_clip_params(seed=0), 'cuda', precision=cprec)
    f = direction.DirectionFinder(G, model, o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'), res, **kw)
    f.delta.copy_(torch.as_tensor(g['delta']).cuda())
    out = f.step(styles.cuda(), lr=0.0)
    ref = torch.as_tensor(g['grad'])[0]
    gr = ((out['grad'].cpu() - ref).norm() / ref.norm()).item()
    rows = ((out['grad'].cpu() - ref).norm(dim=1) / ref.norm(dim=1)).tolist()
    lr = abs(out['clip_loss'].item() - float(g['clip_loss'])) / abs(float(g['clip_loss']))
    print(f'{tag} clip={cprec}: clip-loss rel {lr:.2e} grad rel-l2 {gr:.3e} rows ' + ' '.join(f'{r:.1e}' for r in rows), flush=True)

for cprec in sys.argv[1:] or ['x3p', 'x3', 'x1']:
    g = golden('step64')
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    o_syn.get_temp_shapes(G)
    run('step64 ', G, 64, torch.as_tensor(g['styles']), g, cprec)
    g = golden('config1')
    G = o_syn.make_generator(256, seed=0)
    ws = torch.as_tensor(g['ws'])
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    run('config1', G, 256, S, g, cprec, micro_batch=4)
