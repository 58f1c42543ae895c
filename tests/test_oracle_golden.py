"""CPU: the oracle reproduces the golden vectors that oracle/pin_reference.py recorded from the reference's own code
(torch_utils.ops impl='ref', utils.generate_image, find_direction.unprocess / compute_clip_loss, clip_loss.CLIPLoss)."""
import numpy as np
import pytest
import torch

from oracle import act, conv, direction, fir, synthesis, vit
from cases import CONV_KW, FIR_KW


def T(a):
    return torch.as_tensor(np.asarray(a))


@pytest.mark.parametrize('name', list(FIR_KW))
def test_fir_golden(golden, name):
    g = golden('ops')
    f = T(g[name + '.f']) if name + '.f' in g else None
    assert torch.equal(fir.upfirdn2d(T(g[name + '.x']), f, **FIR_KW[name]), T(g[name + '.y']))


@pytest.mark.parametrize('name', list(act.ACTIVATIONS))
def test_bias_act_golden(golden, name):
    g = golden('ops')
    x, b = T(g['bias_act.x']), T(g['bias_act.b'])
    assert torch.equal(act.bias_act(x, b, act=name), T(g[f'bias_act.{name}.def']))
    assert torch.equal(act.bias_act(x, b, act=name, gain=1.7, clamp=0.9, alpha=0.3), T(g[f'bias_act.{name}.clamp']))


@pytest.mark.parametrize('name', list(CONV_KW) + ['conv_down2', 'conv_1x1_down2'])
def test_conv2d_resample_golden(golden, name):
    g = golden('ops')
    kw = dict(CONV_KW.get(name, {}))
    if name == 'conv_down2':
        kw = dict(down=2, padding=1)
    if name == 'conv_1x1_down2':
        kw = dict(down=2)
    f = fir.setup_filter([1, 3, 3, 1]) if (kw.get('up', 1) > 1 or kw.get('down', 1) > 1) else None
    y = conv.conv2d_resample(T(g[name + '.x']), T(g[name + '.w']), f=f, **kw)
    assert (y - T(g[name + '.y'])).abs().max().item() <= 1e-6      # oneDNN may pick another algorithm between runs


def test_modulated_conv2d_vs_e4e_reference(golden):
    g = golden('ops')
    f4 = fir.setup_filter([1, 3, 3, 1])
    for tag, kw in (('plain', {}), ('up2', dict(up=2, resample_filter=f4, flip_weight=False))):
        y = conv.modulated_conv2d(T(g[f'e4e_{tag}.x']), T(g[f'e4e_{tag}.w']), T(g[f'e4e_{tag}.s']), padding=1, **kw)
        assert (y - T(g[f'e4e_{tag}.y'])).abs().max().item() <= 5e-6


def test_generate_image_golden(golden):
    g = golden('synth64')
    G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    ws = T(g['ws'])
    S, shapes = synthesis.get_styles(G, ws, synthesis.split_ws(G, ws))
    assert (S - T(g['styles'])).abs().max().item() <= 1e-6
    assert [tuple(s) for s in g['temp_shapes']] == shapes
    xs, img = synthesis.generate_image(G, 100, T(g['styles']), shapes)
    assert (img - T(g['img'])).abs().max().item() <= 1e-5
    for i, x in enumerate(xs):
        assert (x - T(g[f'xs{i}'])).abs().max().item() <= 1e-4
    assert (synthesis.generate_image(G, 2, T(g['styles']), shapes)[1] - T(g['img_k2'])).abs().max().item() <= 1e-5


def test_clip_golden(golden):
    g = golden('clip')
    model = vit.CLIP(seed=0)
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        assert (model.encode_image(images) - T(g['image_features'])).abs().max().item() <= 1e-5
        assert (model.encode_text(T(g['tokens'])) - T(g['text_features'])).abs().max().item() <= 1e-5


def test_step_golden(golden):
    g = golden('step64')
    G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    shapes = synthesis.get_temp_shapes(G)
    loss_fn = direction.CLIPLoss(vit.CLIP(seed=0), vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    o = direction.direction_step(G, shapes, loss_fn, T(g['styles']), T(g['delta']), 100)
    assert abs(o['loss'].item() - float(g['loss'])) <= 1e-6
    assert ((o['grad'] - T(g['grad'])).norm() / T(g['grad']).norm()).item() <= 1e-4
    assert (direction.unprocess(T(g['original_img']))[:1] - T(g['unprocessed'])).abs().max().item() <= 1e-5


def test_clip_b16_golden(golden):
    """ViT-B/16 (197 tokens), the second tower of clip_type='double' (clip_loss.py:12-13)."""
    g = golden('clip_b16')
    model = vit.CLIP(seed=int(golden('step64_double')['b16_seed']), cfg=vit.VIT_B16)
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        assert (model.encode_image(images) - T(g['image_features'])).abs().max().item() <= 1e-5
        assert (model.encode_text(T(g['tokens'])) - T(g['text_features'])).abs().max().item() <= 1e-5


def test_step_double_golden(golden):
    """clip_type='double': loss32 + 0.5 * loss16 as the reference's init_clip_loss / compute_clip_loss computed it
    (find_direction.py:117-119,160-164); styles and delta are those of step64.npz."""
    g, gd = golden('step64'), golden('step64_double')
    G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    shapes = synthesis.get_temp_shapes(G)
    pos, neg = vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg')
    loss_fn = direction.DoubleCLIPLoss(direction.CLIPLoss(vit.CLIP(seed=0), pos, neg),
                                       direction.CLIPLoss(vit.CLIP(seed=int(gd['b16_seed']), cfg=vit.VIT_B16), pos, neg))
    o = direction.direction_step(G, shapes, loss_fn, T(g['styles']), T(g['delta']), 100)
    assert abs(o['loss'].item() - float(gd['loss'])) <= 1e-6
    assert abs(o['clip_loss'].item() - float(gd['clip_loss'])) <= 1e-6
    assert ((o['grad'] - T(gd['grad'])).norm() / T(gd['grad']).norm()).item() <= 1e-4


@pytest.mark.parametrize('kind', ['nada', 'nada_global'])
def test_step_nada_golden(golden, kind):
    """clip_loss_type 'nada' / 'nada_global': the reference's real clip_loss_nada.CLIPLoss through init_clip_loss / compute_clip_loss
    (find_direction.py:101-114,150-158; tests/golden/step64_nada.npz) vs oracle.direction.CLIPLossNADA; styles and delta of step64.npz."""
    g, gn = golden('step64'), golden('step64_nada')
    G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    shapes = synthesis.get_temp_shapes(G)
    kw = dict(lambda_direction=0.0, lambda_global=1.0) if kind == 'nada_global' else {}
    loss_fn = direction.CLIPLossNADA(vit.CLIP(seed=0), vit.synthetic_tokenize, **kw)
    o = direction.direction_step(G, shapes, loss_fn, T(g['styles']), T(g['delta']), 100, nada_prompts=(str(gn['neg_text']), str(gn['pos_text'])))
    assert abs(o['loss'].item() - float(gn[kind + '.loss'])) <= 1e-6
    assert ((o['grad'] - T(gn[kind + '.grad'])).norm() / T(gn[kind + '.grad']).norm()).item() <= 1e-4
    assert (direction.nada_preprocess(o['original_img'])[:1] - T(gn['preprocessed'])).abs().max().item() <= 1e-5


def test_idloss_golden(golden):
    """oracle.idloss (IR-SE50 + IDLoss restatement) against the reference's real id_loss.IDLoss / model_irse.Backbone outputs
    (tests/golden/idloss.npz): features, loss, image gradient at 256 px, and the loss at 512 px (exercises the pool to 256)."""
    from oracle import idloss
    g = golden('idloss')
    p = idloss.random_irse50_params(seed=0)
    y, y_hat = T(g['y']), T(g['y_hat']).requires_grad_(True)
    with torch.no_grad():
        assert (idloss.extract_feats(p, y) - T(g['feats_y'])).abs().max().item() <= 1e-6
        assert (idloss.face_crop(y)[:1] - T(g['crop112'])).abs().max().item() <= 1e-6
    loss = idloss.id_loss(p, y_hat, y)
    grad, = torch.autograd.grad(loss, y_hat)
    assert abs(loss.item() - float(g['loss'])) <= 1e-6
    assert ((grad - T(g['grad'])).norm() / T(g['grad']).norm()).item() <= 1e-4


def test_mapper_golden(golden):
    """oracle.mapper (latent_mappers.Mapper restatement + one train_latent_mapper.py:150-176 loss) against the reference's REAL Mapper and
    REAL find_direction.compute_loss (tests/golden/mapper64.npz): delta, loss terms, parameter gradients."""
    from oracle import idloss, mapper
    g, gs = golden('mapper64'), golden('synth64')
    p = mapper.random_mapper_params(seed=3)
    with torch.no_grad():
        assert (mapper.mapper_forward(p, T(g['x'])) - T(g['delta'])).abs().max().item() <= 1e-6
    G = synthesis.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    shapes = synthesis.get_temp_shapes(G)
    S = T(gs['styles'])
    assert torch.equal(S[:, direction.S_TRAINABLE_ROWS], T(g['x']))
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss_fn = direction.CLIPLoss(vit.CLIP(seed=0), vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    o = mapper.mapper_step_loss(G, shapes, loss_fn, pr, S[:2], 100, id_params=idloss.random_irse50_params(seed=0), identity_loss_coef=0.6)
    for k in ('loss', 'clip_loss', 'identity_loss', 'l2_loss'):
        assert abs(o[k].item() - float(g[k])) <= 1e-6, k
    grads = dict(zip(pr, torch.autograd.grad(o['loss'], list(pr.values()))))
    for k, gr in grads.items():
        assert abs(gr.norm().item() - float(g['gradnorm.' + k])) <= 1e-4 * float(g['gradnorm.' + k]), k
        if 'grad.' + k in g:
            assert ((gr - T(g['grad.' + k])).norm() / T(g['grad.' + k]).norm()).item() <= 1e-4, k
