"""GPU parity of the CLIP ViT-B/32 kernels, ``unprocess`` and the directional loss against the golden vectors written by
oracle/pin_reference.py (tests/golden/clip.npz, step64.npz) and the CPU oracle's autograd."""
import pytest
import torch

from oracle import direction as o_dir
from oracle import vit as o_vit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def params():
    return o_vit.random_clip_params(seed=0)


@pytest.fixture(scope='module')
def model(params):
    from stylemc_b200 import clip
    return clip.CLIPModel(params, 'cuda', precision='x3p')


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


def test_encode_golden(golden, model):
    g = golden('clip')
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    ei = model.encode_image(images.cuda())
    et = model.encode_text(torch.as_tensor(g['tokens']).cuda())
    ri, rt = torch.as_tensor(g['image_features']), torch.as_tensor(g['text_features'])
    print('encode_image rel-l2', rel(ei, ri), 'max-abs', (ei.cpu() - ri).abs().max().item(), ' encode_text rel-l2', rel(et, rt))
    assert rel(ei, ri) <= 1e-4 and rel(et, rt) <= 1e-4


def test_encode_image_x1(golden, params):
    from stylemc_b200 import clip
    g = golden('clip')
    m1 = clip.CLIPModel(params, 'cuda', precision='x1')
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    ei = m1.encode_image(images.cuda())
    r = rel(ei, torch.as_tensor(g['image_features']))
    print('x1 encode_image rel-l2', r)
    assert r <= 5e-3


def test_encode_image_input_gradient(model, params):
    oracle = o_vit.CLIP(params)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 224, 224, generator=gen)
    d = torch.randn(3, 512, generator=gen) * 1e-3            # small on purpose: exercises the loss scaling
    xr = x.clone().requires_grad_(True)
    oracle.encode_image(xr).backward(d)
    xc = x.cuda().requires_grad_(True)
    model.encode_image(xc).backward(d.cuda())
    r = rel(xc.grad, xr.grad)
    print('d encode_image / d pixels rel-l2', r)
    assert r <= 1e-3


@pytest.mark.parametrize('res', [64, 256, 512, 1024])
def test_unprocess_fwd_bwd(res):
    from stylemc_b200 import resample
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, res, res, generator=gen) * 0.7      # ~15 % of the pixels hit the clamp(0, 255)
    gy = torch.randn(2, 3, 224, 224, generator=gen)
    xr = x.clone().requires_grad_(True)
    yr = o_dir.unprocess(xr)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    y = resample.unprocess(xc)
    y.backward(gy.cuda())
    # ATen computes the antialias tap weights in float32 (centre = scale * (i + 0.5) carries ~1e-5 absolute error at i ~ 200),
    # so the float32 reference itself is only good to ~1e-4 after /255/std; against the float64 oracle the kernel is tight.
    y64 = o_dir.unprocess(x.double())
    assert (y.detach().cpu().double() - y64).abs().max().item() <= 5e-6
    assert (y.detach().cpu() - yr.detach()).abs().max().item() <= 2e-4
    x64 = x.double().requires_grad_(True)
    o_dir.unprocess(x64).backward(gy.double())
    assert (xc.grad.cpu().double() - x64.grad).abs().max().item() <= 2e-6 * max(1.0, x64.grad.abs().max().item())
    if res <= 256:   # ATen's float32 CPU backward is itself 18 % off (max-abs) at 1024 -> 224; it is fine at 256 and below
        assert (xc.grad.cpu() - xr.grad).abs().max().item() <= 2e-4 * max(1.0, xr.grad.abs().max().item())


def test_unprocess_golden(golden):
    from stylemc_b200 import resample
    g = golden('step64')
    y = resample.unprocess(torch.as_tensor(g['original_img']).cuda())
    assert (y[:1].cpu() - torch.as_tensor(g['unprocessed'])).abs().max().item() <= 2e-4


def test_clip_loss_kernel_and_drop_in(model, params):
    from stylemc_b200 import direction
    oracle = o_vit.CLIP(params)
    pos, neg = o_vit.synthetic_tokens('pos'), o_vit.synthetic_tokens('neg')
    o_loss = o_dir.CLIPLoss(oracle, pos, neg)
    loss = direction.CLIPLoss(model, pos, neg)
    assert rel(loss.text_features, o_loss.text_features) <= 1e-4
    gen = torch.Generator().manual_seed(7)
    src, tgt = torch.randn(3, 3, 224, 224, generator=gen), torch.randn(3, 3, 224, 224, generator=gen)
    tr = tgt.clone().requires_grad_(True)
    lr_ = o_loss(src, tr)
    lr_.backward()
    tc = tgt.cuda().requires_grad_(True)
    lc = loss(src.cuda(), tc)                      # drop-in: torch autograd over encode_image
    lc.backward()
    print('clip loss', lc.item(), lr_.item(), 'grad rel', rel(tc.grad, tr.grad))
    assert abs(lc.item() - lr_.item()) <= 1e-3 * abs(lr_.item())
    assert rel(tc.grad, tr.grad) <= 1e-3
    # fused kernel: same loss and the same embedding gradient
    with torch.no_grad():
        es, et = model.encode_image(src.cuda()), model.encode_image(tgt.cuda())
    etr = et.detach().cpu().clone().requires_grad_(True)
    e = etr - es.cpu()
    e = e / e.norm(dim=1, keepdim=True)
    ref = (3 - torch.nn.functional.cosine_similarity(e, loss.text_features.cpu()).sum()) / 3
    ref.backward()
    part, d_t, gs = loss.loss_and_grad(es, et, 1.0, 1.0 / 3)
    assert abs((1.0 + part.item()) - ref.item()) <= 1e-5
    assert rel(d_t / gs, etr.grad) <= 1e-5
    assert 32.0 <= (d_t.abs().max()).item() < 64.0


@pytest.mark.parametrize('kind', ['nada', 'nada_global'])
def test_step_nada_losses_64px_golden(golden, params, kind):
    """clip_loss_type 'nada' / 'nada_global' (find_direction.py:101-114,150-158) against the reference's real clip_loss_nada.CLIPLoss driven by
    its own init_clip_loss / compute_clip_loss (tests/golden/step64_nada.npz, oracle/pin_reference.py::pin_step_nada): loss and delta-S
    gradient of one step on the 64-px network, BASELINE tolerances; and the preprocessing (no clamp, (x + 1) / 2) against the reference's
    transform pipeline."""
    from oracle import direction as o_dir
    from oracle import synthesis as o_syn
    from stylemc_b200 import clip, direction, resample
    g, g64 = golden('step64_nada'), golden('step64')
    pos_text, neg_text = str(g['pos_text']), str(g['neg_text'])
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    o_syn.get_temp_shapes(G)
    model = clip.CLIPModel(params, 'cuda', precision='x3p')
    if kind == 'nada':
        pos = o_vit.synthetic_tokenize(direction.nada_template_texts(pos_text))
        neg = o_vit.synthetic_tokenize(direction.nada_template_texts(neg_text))
        assert direction.NADA_TEMPLATES == o_dir.NADA_TEMPLATES and pos.shape == (27, 77)
    else:
        pos, neg = o_vit.synthetic_tokenize([f'a {pos_text}']), None
    f = direction.DirectionFinder(G, model, pos, neg, 64, clip_loss_type=kind)
    styles = torch.as_tensor(g64['styles']).cuda()
    f.delta.copy_(torch.as_tensor(g64['delta']).cuda())
    out = f.step(styles, lr=0.0)
    ref_grad = torch.as_tensor(g[kind + '.grad'])[0]
    loss_rel = abs(out['loss'].item() - float(g[kind + '.loss'])) / abs(float(g[kind + '.loss']))
    clip_rel = abs(out['clip_loss'].item() - float(g[kind + '.clip_loss'])) / abs(float(g[kind + '.clip_loss']))
    grad_rel = ((out['grad'].cpu() - ref_grad).norm() / ref_grad.norm()).item()
    print(f'{kind}: loss {out["loss"].item():.6f} ref {float(g[kind + ".loss"]):.6f} rel {loss_rel:.2e}; clip rel {clip_rel:.2e}; grad rel-l2 {grad_rel:.3e}')
    assert loss_rel <= 1e-3 and clip_rel <= 1e-3 and grad_rel <= 1e-3
    # the differentiable drop-in (autograd through resample.nada_preprocess and CLIPModel.encode_image) gives the same loss
    _, img, _ = f.engine.forward(styles + f.direction(), until_k=100)
    _, orig, _ = f.engine.forward(styles, until_k=100)
    drop_in = f.loss_fn(orig, img).item()
    assert abs(drop_in - out['clip_loss'].item()) <= 1e-5 * abs(drop_in)
    pre = resample.nada_preprocess(orig[:1]).cpu()
    assert (pre - torch.as_tensor(g['preprocessed'])).abs().max().item() <= 2e-4
