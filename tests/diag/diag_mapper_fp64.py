"""CPU diagnostic: the fp32 reference gradient of the mapper step (tests/golden/mapper64.npz) against a float64 run of the oracle.  Run from the repo root."""
import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import direction, idloss, mapper, synthesis, vit
torch.set_num_threads(16)
g = np.load('tests/golden/mapper64.npz'); gs = np.load('tests/golden/synth64.npz')
T = torch.as_tensor
def run(dt, coef=0.6):
    G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    if dt == torch.float64: G = G.double()
    shapes = synthesis.get_temp_shapes(G)
    S = T(gs['styles'])[:2].to(dt)
    p = {k: v.to(dt).requires_grad_(True) for k, v in mapper.random_mapper_params(seed=3).items()}
    m = vit.CLIP(seed=0) if dt == torch.float32 else vit.CLIP(params=vit.CLIP(seed=0).p, cfg=vit.VIT_B32, dtype=dt)
    loss_fn = direction.CLIPLoss(m, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    idp = {k: v.to(dt) if v.is_floating_point() else v for k, v in idloss.random_irse50_params(seed=0).items()}
    o = mapper.mapper_step_loss(G, shapes, loss_fn, p, S, 100, id_params=idp, identity_loss_coef=coef)
    grads = dict(zip(p, torch.autograd.grad(o['loss'], list(p.values()))))
    return o, grads
o32, g32 = run(torch.float32)
o64, g64 = run(torch.float64)
print('loss', o32['loss'].item(), o64['loss'].item())
for k in g32:
    if 'grad.' + k in g:
        ref = T(g['grad.' + k])
        print(k, 'golden vs fp64 %.2e   oracle32 vs fp64 %.2e' % (((ref.double() - g64[k]).norm() / g64[k].norm()).item(), ((g32[k].double() - g64[k]).norm() / g64[k].norm()).item()))
np.savez_compressed('/tmp/mapper_grad_fp64.npz', **{k: v.numpy() for k, v in g64.items()})
