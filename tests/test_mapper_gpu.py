"""GPU parity of the latent mapper (stylemc_b200.mapper: latent_mappers.Mapper + one train_latent_mapper.py step on repo kernels) against the
reference's REAL Mapper / find_direction.compute_loss run on the CPU (tests/golden/mapper64.npz, oracle/pin_reference.py::pin_mapper).
Tolerances: BASELINE's (loss and gradients <= 1e-3 relative)."""
import math

import pytest
import torch

from oracle import idloss as o_id
from oracle import mapper as o_map
from oracle import synthesis as o_syn
from oracle import vit as o_vit

pytestmark = pytest.mark.gpu


def T(a):
    return torch.as_tensor(a)


def test_pixelnorm_and_adam_vs_torch():
    from stylemc_b200 import _lib, mapper
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 4, 512, generator=g)
    xc, xr = x.cuda().requires_grad_(True), x.clone().requires_grad_(True)
    y = mapper._PixelNormFn.apply(xc)
    yr = xr * torch.rsqrt(torch.mean(xr ** 2, dim=1, keepdim=True) + 1e-8)
    dy = torch.randn(yr.shape, generator=g)
    y.backward(dy.cuda())
    yr.backward(dy)
    assert (y.detach().cpu() - yr.detach()).abs().max().item() <= 2e-6
    assert (xc.grad.cpu() - xr.grad).abs().max().item() <= 1e-5
    # Adam: three steps against torch.optim.Adam (train_latent_mapper.py:131)
    p0 = torch.randn(1000, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, betas=(0.9, 0.999))
    pc, m, v = p0.cuda(), torch.zeros(1000, device='cuda'), torch.zeros(1000, device='cuda')
    for t in range(1, 4):
        gr = torch.randn(1000, generator=g)
        pr.grad = gr.clone()
        opt.step()
        _lib.call('smc_adam_step', _lib.ptr(pc), _lib.ptr(gr.cuda()), _lib.ptr(m), _lib.ptr(v), 1000, 1e-3, 0.9, 0.999, 1e-8, 1.0 - 0.9 ** t,
                  math.sqrt(1.0 - 0.999 ** t), _lib.stream())
        assert (pc.cpu() - pr.detach()).abs().max().item() <= 1e-6


def test_mapper_forward_and_param_grads_vs_oracle(golden):
    """Mapper.forward against the reference's delta (golden) and d(sum(delta * r))/d(parameters) against the oracle's autograd."""
    from stylemc_b200 import mapper
    g = golden('mapper64')
    p = o_map.random_mapper_params(seed=3)
    m = mapper.Mapper(neg_slope=0.01)
    m.load_state_dict(p)
    x = T(g['x'])
    delta = m(x.cuda())
    err = (delta.detach().cpu() - T(g['delta'])).abs().max().item()
    print('mapper forward max-abs err', err)
    # outputs are O(1) (LayerNorm then LeakyReLU).  The bar is tight on purpose: delta is added to S, and a 5e-6 error in it moves the
    # synthesis gradient by 8e-4 (float64 oracle) -- the linears accumulate in float64 (smc_matmul_nt_f64acc)
    assert err <= 2e-6
    r = torch.randn(delta.shape, generator=torch.Generator().manual_seed(6))
    delta.backward(r.cuda())
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = dict(zip(pr, torch.autograd.grad((o_map.mapper_forward(pr, x) * r).sum(), list(pr.values()))))
    worst = max(((m.params[k].grad.cpu() - ref[k]).norm() / ref[k].norm()).item() for k in ref)
    print('worst parameter-gradient rel-l2', worst)
    assert worst <= 1e-3
    sd = m.state_dict()
    assert set(sd) == set(p) and all(torch.equal(sd[k].cpu(), p[k]) for k in p)


def test_mapper_step_golden(golden):
    """One train_latent_mapper.py:150-176 step (CLIP + 0.6 identity + 0.1 L2, per-image delta) against the reference's compute_loss + autograd."""
    from stylemc_b200 import clip, idloss, mapper
    g, gs = golden('mapper64'), golden('synth64')
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    S = T(gs['styles'])[:2]
    m = mapper.Mapper(neg_slope=0.01)
    p = o_map.random_mapper_params(seed=3)
    m.load_state_dict(p)
    id_mod = idloss.IDLoss(o_id.random_irse50_params(seed=0), 'cuda')
    tr = mapper.MapperTrainer(G, clip.CLIPModel(o_vit.random_clip_params(seed=0), 'cuda'), o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg'),
                              64, m, id_loss=id_mod, identity_loss_coef=0.6, micro_batch=2)
    out, grads = tr.loss_and_grads(S.cuda())
    for k in ('loss', 'clip_loss', 'identity_loss', 'l2_loss'):
        rel = abs(out[k].item() - float(g[k])) / abs(float(g[k]))
        print(f'{k}: {out[k].item():.6f} ref {float(g[k]):.6f} rel {rel:.2e}')
        assert rel <= 1e-3, k
    worst = 0.0
    for k, gr in grads.items():
        n_ref = float(g['gradnorm.' + k])
        assert abs(gr.norm().item() - n_ref) <= 2e-3 * n_ref, k
        if 'grad.' + k in g:
            ref = T(g['grad.' + k])
            err = ((gr.cpu() - ref).norm() / ref.norm()).item()
            print(f'  {k}: rel-l2 {err:.2e}')
            worst = max(worst, err)
    print('worst parameter-gradient rel-l2 vs the reference', worst)
    assert worst <= 1e-3
    # one Adam step moves every parameter by ~lr (first step: |update| = lr * g / (|g| + eps))
    before = {k: v.detach().clone() for k, v in m.params.items()}
    tr.step(S.cuda())
    for k, v in m.params.items():
        d = (v.detach() - before[k]).abs().max().item()
        assert 0.0 < d <= 1.001e-3, (k, d)


def test_per_sample_gradient_sums_to_the_shared_one(monkeypatch):
    """SynthesisEngine.backward(per_sample=True): the per-image style gradients sum to the shared-delta gradient (find_direction.py:307-308):
    to rounding when both passes use the same gradient planes, to the two-term split's error with the per-sample
    default (hi + lo gradient planes: no batch sum for a single plane's rounding to average out in)."""
    from stylemc_b200 import clip, direction
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    ws = torch.randn(3, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(21))
    S, _ = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    f = direction.DirectionFinder(G, clip.CLIPModel(o_vit.random_clip_params(seed=0), 'cuda'), o_vit.synthetic_tokens('pos'),
                                  o_vit.synthetic_tokens('neg'), 64, micro_batch=2)
    f.delta.copy_(0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(22)).cuda())
    g_sum, _ = f.loss_and_grad(S.cuda(), 3)
    g_each, _ = f.loss_and_grad(S.cuda(), 3, per_sample=True)
    assert g_each.shape == (3, 8, 512)
    assert ((g_each.sum(0) - g_sum).norm() / g_sum.norm()).item() <= 1e-3
    monkeypatch.setenv('STYLEMC_MAPPER_GRAD_LO', '0')
    g_each, _ = f.loss_and_grad(S.cuda(), 3, per_sample=True)
    assert ((g_each.sum(0) - g_sum).norm() / g_sum.norm()).item() <= 1e-5
