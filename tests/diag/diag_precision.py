"""Dev diagnostic: config-1 step gradient error for each engine precision (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import synthesis as o_syn, vit as o_vit
from stylemc_b200 import clip, direction

g = np.load('tests/golden/config1.npz')
g64 = torch.as_tensor(np.load('tests/golden/config1_grad_fp64.npy'))[0].float()
G = o_syn.make_generator(256, seed=0)
ws = torch.as_tensor(g['ws'])
S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
model = clip.CLIPModel(o_vit.random_clip_params(seed=0), 'cuda', precision='x3p')
ref = torch.as_tensor(g['grad'])[0]
for prec in sys.argv[1:] or ['x1', 'mixed', 'x3', 'x3p']:
    f = direction.DirectionFinder(G, model, o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'), 256, precision=prec, micro_batch=4)
    f.delta.copy_(torch.as_tensor(g['delta']).cuda())
    out = f.step(S.cuda(), lr=0.0)
    gr = out['grad'].cpu()
    print(f'{prec}: loss rel {abs(out["loss"].item()-float(g["loss"]))/float(g["loss"]):.2e}  grad rel vs fp32 ref {((gr-ref).norm()/ref.norm()).item():.3e}  vs fp64 {((gr-g64).norm()/g64.norm()).item():.3e}')
    for i, r in enumerate(direction.S_TRAINABLE_SPACE_CHANNELS):
        print(f'   row {r}: {((gr[i]-g64[i]).norm()/g64[i].norm()).item():.3e}')
