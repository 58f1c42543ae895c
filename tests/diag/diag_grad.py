"""Dev diagnostic (GPU box): where does the config-1 delta-S gradient error come from?  Splits it into the CLIP/unprocess backward
(gradient w.r.t. the image) and the synthesis backward, using the CPU oracle's own image gradient as the hand-over."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import synthesis as o_syn, vit as o_vit, direction as o_dir
from stylemc_b200 import clip, direction, resample

torch.set_num_threads(os.cpu_count())
g = np.load('tests/golden/config1.npz')
g64 = torch.as_tensor(np.load('tests/golden/config1_grad_fp64.npy'))[0].float()
G = o_syn.make_generator(256, seed=0)
ws = torch.as_tensor(g['ws'])
S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
delta0 = torch.as_tensor(g['delta'])
rows = o_dir.S_TRAINABLE_ROWS

# ---- oracle (fp32 CPU) with the image gradient exposed
params = o_vit.random_clip_params(seed=0)
loss_fn = o_dir.CLIPLoss(o_vit.CLIP(params), o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'))
delta = delta0.clone().requires_grad_(True)
dirn = torch.zeros(1, 26, 512).index_put((torch.tensor([0]).view(1, 1), torch.tensor(rows).view(1, -1)), delta)
_, img = o_syn.generate_image(G, 6, S + dirn, shapes, 'const')
img.retain_grad()
with torch.no_grad():
    _, orig = o_syn.generate_image(G, 6, S, shapes, 'const')
clip_term = loss_fn(o_dir.unprocess(orig), o_dir.unprocess(img))
clip_term.backward()
g_img_ref = img.grad.clone()
grad_ref = delta.grad[0].clone()          # clip part only (no l2)
print('oracle clip loss', clip_term.item(), ' |g_img|', g_img_ref.norm().item(), ' |grad|', grad_ref.norm().item())

# ---- ours
model = clip.CLIPModel(params, 'cuda', precision='x3p')
f = direction.DirectionFinder(G, model, o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'), 256, precision='x3p', micro_batch=4)
f.delta.copy_(delta0.cuda())
eng = f.engine
s = S.cuda()
s2 = s + f.direction()
_, img_o, saved = eng.forward(s2, f.until_k, 'const', save=True)
_, orig_o, _ = eng.forward(s, f.until_k, 'const', save=False)
print('img err', (img_o.cpu() - img.detach()).abs().max().item(), ' orig err', (orig_o.cpu() - orig).abs().max().item())
u_t = resample.unprocess_fwd(img_o)
u_s = resample.unprocess_fwd(orig_o)
e_s, _ = f.clip.encode_image_fwd(u_s, save=False)
e_t, csaved = f.clip.encode_image_fwd(u_t, save=True)
part, d_t, gscale = f.loss_fn.loss_and_grad(e_s, e_t, 1.0, 1.0 / 4)
g224 = f.clip.encode_image_bwd(csaved, d_t)
g_img_o = resample.unprocess_bwd(g224, img_o, unscale=gscale)
rel = lambda a, b: ((a - b).norm() / b.norm()).item()
print('g_img  ours vs oracle rel-l2:', rel(g_img_o.cpu(), g_img_ref))
grad_full = eng.backward(saved, g_img_o, f.rows, 'const').cpu()
grad_synth = eng.backward(saved, g_img_ref.cuda(), f.rows, 'const').cpu()
print('grad (ours end to end)            vs oracle:', rel(grad_full, grad_ref))
print('grad (oracle g_img -> our synth bwd) vs oracle:', rel(grad_synth, grad_ref))
for i, r in enumerate(rows):
    print(f'   row {r}: e2e {rel(grad_full[i], grad_ref[i]):.3e}   synth-bwd only {rel(grad_synth[i], grad_ref[i]):.3e}')

# oracle: own sensitivity -- the same backward with the image gradient perturbed like ours

print('--- sensitivity (oracle g_img -> our synthesis fwd+bwd) ---')
for acc_k in (512, 256, 128):
    for tgt in (256.0, 4096.0, 16.0):
        eng.acc_k = acc_k
        _, _, sv = eng.forward(s2, f.until_k, 'const', save=True)
        gr = eng.backward(sv, g_img_ref.cuda(), f.rows, 'const', grad_scale_target=tgt).cpu()
        print(f'acc_k {acc_k} gscale target {tgt}: total {rel(gr, grad_ref):.3e}  rows ' + ' '.join(f'{rel(gr[i], grad_ref[i]):.2e}' for i in range(len(rows))))
eng.acc_k = 512

print('--- low-resolution chain length (forward convs) ---')
for lowmax in (0, 16, 32, 64):
    for lowk in (64, 128):
        eng.acc_k_lowres, eng.acc_k_lowres_max = lowk, lowmax
        _, _, sv = eng.forward(s2, f.until_k, 'const', save=True)
        gr = eng.backward(sv, g_img_ref.cuda(), f.rows, 'const').cpu()
        print(f'lowres<= {lowmax} acc_k {lowk}: total {rel(gr, grad_ref):.3e}  rows ' + ' '.join(f'{rel(gr[i], grad_ref[i]):.2e}' for i in range(len(rows))))
