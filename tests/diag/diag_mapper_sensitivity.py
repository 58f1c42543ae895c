"""CPU diagnostic (float64 oracle): how much the per-sample synthesis gradient moves when the mapper's delta is perturbed.  Run from the repo root."""
import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import direction, idloss, mapper, synthesis, vit
import torch.nn.functional as F
torch.set_num_threads(16)
gs = np.load('tests/golden/synth64.npz')
T = torch.as_tensor
dt = torch.float64
G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512).double()
shapes = synthesis.get_temp_shapes(G)
S = T(gs['styles'])[:2].to(dt)
rows = direction.S_TRAINABLE_ROWS
p = {k: v.to(dt) for k, v in mapper.random_mapper_params(seed=3).items()}
m = vit.CLIP(params=vit.CLIP(seed=0).p, cfg=vit.VIT_B32, dtype=dt)
loss_fn = direction.CLIPLoss(m, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
idp = {k: v.to(dt) if v.is_floating_point() else v for k, v in idloss.random_irse50_params(seed=0).items()}
delta0 = mapper.mapper_forward(p, S[:, rows]).detach()
def grad_wrt_delta(delta, coef=0.6):
    delta = delta.clone().requires_grad_(True)
    s2 = S.clone(); s2[:, rows] = s2[:, rows] + delta
    _, img = synthesis.generate_image(G, 100, s2, shapes)
    with torch.no_grad(): _, orig = synthesis.generate_image(G, 100, S, shapes)
    l = loss_fn(direction.unprocess(orig), direction.unprocess(img)) + (coef * idloss.id_loss(idp, img, orig) if coef else 0.0)
    return torch.autograd.grad(l, delta)[0]
g0 = grad_wrt_delta(delta0)
for eps in (1e-7, 1e-6, 5e-6):
    g1 = grad_wrt_delta(delta0 + eps * torch.randn(delta0.shape, dtype=dt, generator=torch.Generator().manual_seed(1)))
    print('delta perturbed by', eps, '-> gradient changes by rel-l2 %.2e' % ((g1 - g0).norm() / g0.norm()).item())
print('per-sample norms', g0[0].norm().item(), g0[1].norm().item(), 'per-row norms sample0', [round(g0[0, i].norm().item(), 5) for i in range(8)])
