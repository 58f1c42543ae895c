"""Dev diagnostic (GPU box): how much precision do the HIGH-resolution blocks need?  1024-px config-f network, 2 seeds, CPU oracle
(fp32) as reference; our step with split precision everywhere vs fp16 operands (x1) from a given resolution up."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import synthesis as o_syn, vit as o_vit, direction as o_dir
from stylemc_b200 import clip, direction

torch.set_num_threads(os.cpu_count())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
G = o_syn.make_generator(1024, seed=0)
ws = torch.randn(n, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(3))
S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
delta0 = 0.05 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(7))
params = o_vit.random_clip_params(seed=0)
loss_fn = o_dir.CLIPLoss(o_vit.CLIP(params), o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'))
t0 = time.time()
ref = o_dir.direction_step(G, shapes, loss_fn, S, delta0, 8)
print(f'oracle: {time.time() - t0:.1f} s, loss {ref["loss"].item():.6f} |grad| {ref["grad"].norm().item():.4e}', flush=True)
grad_ref = ref['grad'][0]
l2 = 2.0 * 0.1 / delta0.numel() * delta0[0]            # the oracle's gradient includes the L2 term; ours is compared on the same footing

model = clip.CLIPModel(params, 'cuda', precision='x3p')
rel = lambda a, b: ((a - b).norm() / b.norm()).item()
for x1_from in (0, 1024, 512, 256, 128):
    f = direction.DirectionFinder(G, model, o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'), 1024, precision='x3p', micro_batch=n)
    f.delta.copy_(delta0.cuda())
    eng = f.engine
    if x1_from:
        eng._prec = (lambda res, t=x1_from: 'x1' if res >= t else 'x3')
    out = f.step(S.cuda(), lr=0.0)
    g = out['grad'].cpu()
    _, img, _ = eng.forward(S.cuda() + f.direction(), until_k=f.until_k)
    print(f'x1 from {x1_from or "never"}: loss rel {abs(out["loss"].item() - ref["loss"].item()) / abs(ref["loss"].item()):.2e}  img max-abs {(img.cpu() - ref["img"]).abs().max().item():.2e}  '
          f'grad rel-l2 {rel(g, grad_ref):.3e}  rows ' + ' '.join(f'{rel(g[i], grad_ref[i]):.1e}' for i in range(8)), flush=True)
