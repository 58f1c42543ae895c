"""GPU parity of the identity loss (stylemc_b200.idloss: IR-SE50 on repo kernels) against the reference's real id_loss.IDLoss /
model_irse.Backbone run on the CPU (tests/golden/idloss.npz, oracle/pin_reference.py::pin_idloss) and against the CPU oracle inside a whole
find_direction step.  Tolerances: BASELINE's (loss and gradient <= 1e-3 relative); features are compared as unit vectors."""
import pytest
import torch

from oracle import direction as o_dir
from oracle import idloss as o_id
from oracle import synthesis as o_syn
from oracle import vit as o_vit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def id_mod():
    from stylemc_b200 import idloss
    return idloss.IDLoss(o_id.random_irse50_params(seed=0), 'cuda')


def test_pools_and_prelu_vs_torch():
    from stylemc_b200 import idloss
    g = torch.Generator().manual_seed(3)
    for res in (256, 300, 1024):
        x = torch.randn(2, 3, res, res, generator=g)
        xc = x.cuda().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        y, yr = idloss.face_crop(xc), o_id.face_crop(xr)
        assert y.shape == yr.shape == (2, 3, 112, 112)
        assert (y.detach().cpu() - yr.detach()).abs().max().item() <= 2e-6
        dy = torch.randn(yr.shape, generator=g)
        y.backward(dy.cuda())
        yr.backward(dy)
        assert (xc.grad.cpu() - xr.grad).abs().max().item() <= 1e-6 * max(1.0, xr.grad.abs().max().item())
    x = torch.randn(3, 40, 9, 7, generator=g)
    a = torch.rand(40, generator=g)
    xc = x.cuda().requires_grad_(True)
    y = idloss._PreluFn.apply(xc, a.cuda())
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.prelu(xr, a)
    dy = torch.randn(yr.shape, generator=g)
    y.backward(dy.cuda())
    yr.backward(dy)
    assert torch.equal(y.detach().cpu(), yr.detach()) and torch.equal(xc.grad.cpu(), xr.grad)


def test_idloss_golden(golden, id_mod):
    g = golden('idloss')
    y, y_hat = torch.as_tensor(g['y']).cuda(), torch.as_tensor(g['y_hat']).cuda()
    from stylemc_b200 import idloss
    assert (idloss.face_crop(y)[:1].cpu() - torch.as_tensor(g['crop112'])).abs().max().item() <= 2e-6
    f = id_mod.extract_feats(y).cpu()
    ref = torch.as_tensor(g['feats_y'])
    print('feature max-abs err', (f - ref).abs().max().item(), 'cos', (f * ref).sum(1).tolist())
    assert (f - ref).abs().max().item() <= 1e-4                               # unit vectors of 512 entries (~0.04 each)
    part, grad = id_mod.loss_and_grad(y_hat, y)
    loss = 1.0 + part.item()
    loss_rel = abs(loss - float(g['loss'])) / float(g['loss'])
    ref_grad = torch.as_tensor(g['grad'])
    grad_rel = ((grad.cpu() - ref_grad).norm() / ref_grad.norm()).item()
    print(f'id loss {loss:.7f} ref {float(g["loss"]):.7f} rel {loss_rel:.2e}; image-gradient rel-l2 {grad_rel:.3e}')
    assert loss_rel <= 1e-3 and grad_rel <= 1e-3
    # the reference call surface: (loss, sim_improvement) = id_loss(y_hat, y), differentiable w.r.t. y_hat
    yh = y_hat.clone().requires_grad_(True)
    l, _ = id_mod(yh, y)
    l.backward()
    assert abs(l.item() - loss) <= 1e-6 and ((yh.grad.cpu() - ref_grad).norm() / ref_grad.norm()).item() <= 1e-3


def test_step_with_identity_term_vs_oracle(id_mod):
    """One find_direction step with the identity term switched on (identity_loss_coef 0.6, the reference CLI's default, find_direction.py:225)
    on the 64-px network against the CPU oracle: CLIP term + L2 term + 0.6 * id_loss(img, original) (find_direction.py:179-193)."""
    from stylemc_b200 import clip, direction
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    ws = torch.randn(3, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(21))
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    delta = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(22))
    params = o_vit.random_clip_params(seed=0)
    pos, neg = o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg')
    # oracle: the step of oracle.direction plus the identity term on the same images
    d = delta.clone().requires_grad_(True)
    direction_t = torch.zeros(1, 26, 512).index_put((torch.tensor([0]).view(1, 1), torch.tensor(o_dir.S_TRAINABLE_ROWS).view(1, -1)), d)
    _, img = o_syn.generate_image(G, 100, S + direction_t, shapes)
    with torch.no_grad():
        _, orig = o_syn.generate_image(G, 100, S, shapes)
    clip_term = o_dir.CLIPLoss(o_vit.CLIP(params), pos, neg)(o_dir.unprocess(orig), o_dir.unprocess(img))
    id_term = 0.6 * o_id.id_loss(o_id.random_irse50_params(seed=0), img, orig)
    l2 = 0.1 * torch.nn.functional.mse_loss((S + direction_t)[:, o_dir.S_TRAINABLE_ROWS], S[:, o_dir.S_TRAINABLE_ROWS])
    loss = clip_term + id_term + l2
    ref_grad, = torch.autograd.grad(loss, d)
    g_id, = torch.autograd.grad(id_term, d, retain_graph=False) if False else (None,)
    f = direction.DirectionFinder(G, clip.CLIPModel(params, 'cuda'), pos, neg, 64, id_loss=id_mod, identity_loss_coef=0.6, micro_batch=2)
    f.delta.copy_(delta.cuda())
    out = f.step(S.cuda(), lr=0.0)
    grad_rel = ((out['grad'].cpu() - ref_grad[0]).norm() / ref_grad.norm()).item()
    print(f'loss {out["loss"].item():.6f} ref {loss.item():.6f}; identity {out["identity_loss"].item():.6f} ref {id_term.item():.6f}; grad rel-l2 {grad_rel:.3e}')
    assert abs(out['loss'].item() - loss.item()) <= 1e-3 * abs(loss.item())
    assert abs(out['identity_loss'].item() - id_term.item()) <= 1e-3 * abs(id_term.item())
    assert grad_rel <= 1e-3
