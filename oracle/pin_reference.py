#!/usr/bin/env python
"""Pin the oracle against the REAL reference code and write ``tests/golden/*.npz``.

Runs only where ``/root/reference`` is mounted (the build container); the fixtures it writes are
committed so the GPU box (which has no reference tree) can check both the oracle and the CUDA path.

    python oracle/pin_reference.py            # verify + (re)write fixtures
    python oracle/pin_reference.py --check    # verify only

What runs from the reference, unmodified (imported, never copied):
  torch_utils.ops.upfirdn2d / bias_act / conv2d_resample / fma      (impl='ref' CPU branches)
  utils.get_temp_shapes / get_styles / split_ws / generate_image / block_forward / get_mean_std
  find_direction.unprocess / compute_clip_loss / S_TRAINABLE_SPACE_CHANNELS
  clip_loss.CLIPLoss   (with a stub ``clip`` module: ``clip.load`` -> oracle ViT-B/32, ``clip.tokenize`` ->
                        fixed synthetic tokens; the real package is not installable offline)
What cannot run here: training/networks.py (absent) and openai/CLIP -> restated in oracle.synthesis /
oracle.vit; the latter is cross-checked against transformers.CLIPModel by weight copy.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, REPO)

from oracle import act, conv, direction, fir, idloss, mapper, synthesis, vit  # noqa: E402

GOLD = os.path.join(REPO, 'tests', 'golden')


def import_reference():
    assert os.path.isdir(REF), 'reference tree not mounted'
    sys.path.insert(0, REF)
    stub_clip = types.ModuleType('clip')
    sys.modules['clip'] = stub_clip
    for name in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    os.environ.setdefault('WANDB_MODE', 'disabled')
    from torch_utils.ops import bias_act, conv2d_resample, fma, upfirdn2d
    import utils as ref_utils
    import find_direction as ref_fd
    import clip_loss as ref_cl
    return types.SimpleNamespace(upfirdn2d=upfirdn2d, bias_act=bias_act, conv2d_resample=conv2d_resample, fma=fma,
                                 utils=ref_utils, fd=ref_fd, cl=ref_cl, clip=stub_clip)


def close(a, b, tol, what):
    err = (a.double() - b.double()).abs().max().item()
    ok = err <= tol
    print(f'  [{"ok" if ok else "FAIL"}] {what}: max-abs {err:.3e} (tol {tol:.0e})')
    assert ok, what
    return err


# ------------------------------------------------------------------------------------------------

def pin_ops(R, out):
    print('ops: oracle vs reference torch_utils.ops (impl="ref")')
    g = torch.Generator().manual_seed(0)
    f4 = R.upfirdn2d.setup_filter([1, 3, 3, 1])
    close(fir.setup_filter([1, 3, 3, 1]), f4, 0, 'setup_filter [1,3,3,1]')
    f8 = R.upfirdn2d.setup_filter([1, 2, 3, 4, 4, 3, 2, 1])
    close(fir.setup_filter([1, 2, 3, 4, 4, 3, 2, 1]), f8, 0, 'setup_filter separable 8 taps')
    close(fir.setup_filter([1, 2, 1], gain=3, flip_filter=True), R.upfirdn2d.setup_filter([1, 2, 1], gain=3, flip_filter=True), 0, 'setup_filter gain/flip')

    fir_cases = {
        'fir_conv0': dict(shape=(2, 5, 17, 17), f=f4, kw=dict(padding=[1, 1, 1, 1], gain=4)),
        'fir_up2': dict(shape=(2, 3, 8, 8), f=f4, kw=dict(up=2, padding=[2, 1, 2, 1], gain=4)),
        'fir_down2': dict(shape=(2, 3, 16, 16), f=f4, kw=dict(down=2, padding=[1, 1, 1, 1])),
        'fir_down2_bwd_of_up2': dict(shape=(1, 3, 16, 16), f=f4, kw=dict(down=2, padding=[1, 2, 1, 2], flip_filter=True, gain=4)),
        'fir_crop_flip': dict(shape=(1, 2, 12, 9), f=torch.rand(3, 5, generator=g), kw=dict(padding=[-1, 2, 0, -2], flip_filter=True, gain=0.5)),
        'fir_updown_xy': dict(shape=(1, 2, 7, 6), f=torch.rand(4, 4, generator=g), kw=dict(up=[2, 3], down=[3, 2], padding=[3, 2, 4, 1])),
        'fir_sep8': dict(shape=(1, 2, 20, 20), f=f8, kw=dict(up=2, padding=[4, 3, 4, 3], gain=4)),
        'fir_identity': dict(shape=(1, 2, 5, 5), f=None, kw=dict(up=2, padding=1)),
    }
    for name, c in fir_cases.items():
        x = torch.randn(*c['shape'], generator=g)
        yr = R.upfirdn2d.upfirdn2d(x, c['f'], impl='ref', **c['kw'])
        close(fir.upfirdn2d(x, c['f'], **c['kw']), yr, 0, name)
        out[name + '.x'], out[name + '.y'] = x.numpy(), yr.numpy()
        if c['f'] is not None:
            out[name + '.f'] = c['f'].numpy()
    x = torch.randn(2, 3, 8, 8, generator=g)
    close(fir.upsample2d(x, f4), R.upfirdn2d.upsample2d(x, f4, impl='ref'), 0, 'upsample2d')
    close(fir.downsample2d(x, f4), R.upfirdn2d.downsample2d(x, f4, impl='ref'), 0, 'downsample2d')
    close(fir.filter2d(x, f4), R.upfirdn2d.filter2d(x, f4, impl='ref'), 0, 'filter2d')
    out['upsample2d.x'], out['upsample2d.y'] = x.numpy(), R.upfirdn2d.upsample2d(x, f4, impl='ref').numpy()

    x = torch.randn(3, 6, 5, 4, generator=g) * 3
    b = torch.randn(6, generator=g)
    out['bias_act.x'], out['bias_act.b'] = x.numpy(), b.numpy()
    for name in act.ACTIVATIONS:
        for tag, kw in (('def', {}), ('clamp', dict(gain=1.7, clamp=0.9, alpha=0.3))):
            yr = R.bias_act.bias_act(x, b, act=name, impl='ref', **kw)
            close(act.bias_act(x, b, act=name, **kw), yr, 0, f'bias_act {name} {tag}')
            out[f'bias_act.{name}.{tag}'] = yr.numpy()
        spec, rspec = act.ACTIVATIONS[name], R.bias_act.activation_funcs[name]
        assert (spec.def_alpha, spec.cuda_idx, spec.ref, spec.has_2nd_grad) == (rspec.def_alpha, rspec.cuda_idx, rspec.ref, rspec.has_2nd_grad)
        assert abs(spec.def_gain - rspec.def_gain) < 1e-12
    x2 = torch.randn(4, 7, generator=g)
    b2 = torch.randn(4, generator=g)
    close(act.bias_act(x2, b2, dim=0, act='lrelu'), R.bias_act.bias_act(x2, b2, dim=0, act='lrelu', impl='ref'), 0, 'bias_act dim=0')
    # first-order gradient of the path's two uses
    for kw in (dict(act='lrelu', gain=2 ** 0.5, clamp=256 * 2 ** 0.5), dict(act='linear', clamp=256), dict(act='lrelu', gain=1.5, clamp=1.0)):
        xa = (x * 60).clone().requires_grad_(True)
        xb = (x * 60).clone().requires_grad_(True)
        dy = torch.randn(x.shape, generator=g)
        R.bias_act.bias_act(xa, b, impl='ref', **kw).backward(dy)
        act.bias_act(xb, b, **kw).backward(dy)
        close(xb.grad, xa.grad, 0, f'bias_act grad {kw}')

    conv_cases = {
        'conv_up2': dict(x=(2, 6, 5, 5), w=(4, 6, 3, 3), kw=dict(up=2, padding=1, flip_weight=False)),
        'conv_up2_grouped': dict(x=(1, 6, 5, 5), w=(8, 3, 3, 3), kw=dict(up=2, padding=1, groups=2, flip_weight=False)),
        'conv_plain': dict(x=(2, 6, 8, 8), w=(4, 6, 3, 3), kw=dict(padding=1)),
        'conv_down2': dict(x=(2, 6, 8, 8), w=(4, 6, 3, 3), kw=dict(down=2, padding=1)),
        'conv_1x1': dict(x=(2, 6, 8, 8), w=(3, 6, 1, 1), kw=dict()),
        'conv_1x1_up2': dict(x=(2, 6, 4, 4), w=(3, 6, 1, 1), kw=dict(up=2)),
        'conv_1x1_down2': dict(x=(2, 6, 8, 8), w=(3, 6, 1, 1), kw=dict(down=2)),
    }
    for name, c in conv_cases.items():
        x = torch.randn(*c['x'], generator=g)
        w = torch.randn(*c['w'], generator=g)
        f = f4 if (c['kw'].get('up', 1) > 1 or c['kw'].get('down', 1) > 1) else None
        yr = R.conv2d_resample.conv2d_resample(x, w, f=f, **c['kw'])
        close(conv.conv2d_resample(x, w, f=f, **c['kw']), yr, 0, name)
        out[name + '.x'], out[name + '.w'], out[name + '.y'] = x.numpy(), w.numpy(), yr.numpy()
    a, bb, cc = torch.randn(2, 3, 4, 4, generator=g), torch.randn(2, 3, 1, 1, generator=g), torch.randn(4, 4, generator=g)
    close(conv.fma(a, bb, cc), R.fma.fma(a, bb, cc), 0, 'fma')
    # fma.py:15-58 forward + broadcast-aware backward through the reference's own autograd.Function
    ar, br, cr = (t.clone().requires_grad_(True) for t in (a, bb, cc))
    dy = torch.randn(2, 3, 4, 4, generator=g)
    yr = R.fma.fma(ar, br, cr)
    yr.backward(dy)
    out['fma.a'], out['fma.b'], out['fma.c'], out['fma.dy'], out['fma.y'] = a.numpy(), bb.numpy(), cc.numpy(), dy.numpy(), yr.detach().numpy()
    out['fma.da'], out['fma.db'], out['fma.dc'] = ar.grad.numpy(), br.grad.numpy(), cr.grad.numpy()
    # the conv2d_gradfix entry points (conv2d_gradfix.py:35-43; pass-through to F.conv2d / F.conv_transpose2d on this torch)
    from torch_utils.ops import conv2d_gradfix as ref_gradfix
    xg, wg, bg = torch.randn(2, 32, 9, 9, generator=g), torch.randn(32, 32, 3, 3, generator=g), torch.randn(32, generator=g)
    out['gradfix.x'], out['gradfix.w'], out['gradfix.bias'] = xg.numpy(), wg.numpy(), bg.numpy()
    out['gradfix.conv2d'] = ref_gradfix.conv2d(xg, wg, bias=bg, padding=1).numpy()
    wt = torch.randn(32, 32, 3, 3, generator=g)                       # conv_transpose2d weight: [in, out, kh, kw]
    out['gradfix.wt'] = wt.numpy()
    out['gradfix.conv_transpose2d'] = ref_gradfix.conv_transpose2d(xg, wt, stride=2).numpy()
    xr_ = xg.clone().requires_grad_(True)
    dyc = torch.randn(2, 32, 9, 9, generator=g)
    ref_gradfix.conv2d(xr_, wg, padding=1).backward(dyc)
    out['gradfix.dy'], out['gradfix.dx'] = dyc.numpy(), xr_.grad.numpy()

    # modulated_conv2d: upstream absent; self-consistency fused vs non-fused (7e-7)
    x = torch.randn(3, 8, 6, 6, generator=g)
    w = torch.randn(5, 8, 3, 3, generator=g)
    s = torch.randn(3, 8, generator=g) + 1
    nz = torch.randn(12, 12, generator=g) * 0.1
    for kw in (dict(padding=1), dict(up=2, padding=1, resample_filter=f4, flip_weight=False, noise=nz), dict(demodulate=False)):
        yf = conv.modulated_conv2d(x, w, s, fused_modconv=True, **kw)
        yn = conv.modulated_conv2d(x, w, s, fused_modconv=False, **kw)
        close(yf, yn, 2e-5, f'modulated_conv2d fused vs non-fused {list(kw)}')
    out['modconv.x'], out['modconv.w'], out['modconv.s'], out['modconv.noise'] = x.numpy(), w.numpy(), s.numpy(), nz.numpy()
    out['modconv.y_plain'] = conv.modulated_conv2d(x.double(), w.double(), s.double(), padding=1).float().numpy()
    out['modconv.y_up2'] = conv.modulated_conv2d(x.double(), w.double(), s.double(), up=2, padding=1, resample_filter=f4, flip_weight=False, noise=nz.double()).float().numpy()


def pin_modconv_e4e(R, out):
    """Anchor oracle.conv.modulated_conv2d on the in-tree analogue
    encoder4editing/models/stylegan2/model.py:177-273 (rosinality ModulatedConv2d: scale*W*s,
    rsqrt(sum w^2 + 1e-8), grouped conv / conv_transpose2d(stride 2) + blur), run from the reference
    with its CUDA-only ``op`` package stubbed by the reference's own torch_utils upfirdn2d (impl='ref')."""
    print('modconv: oracle.modulated_conv2d vs reference e4e ModulatedConv2d')
    op = types.ModuleType('encoder4editing.models.stylegan2.op')
    op.upfirdn2d = lambda x, k, up=1, down=1, pad=(0, 0): R.upfirdn2d.upfirdn2d(
        x, k, up=up, down=down, padding=[pad[0], pad[1], pad[0], pad[1]], impl='ref')
    op.FusedLeakyReLU = torch.nn.Identity
    op.fused_leaky_relu = lambda x, b=None, *a, **k: x
    sys.modules['encoder4editing.models.stylegan2.op'] = op
    sys.path.insert(0, REF)
    from encoder4editing.models.stylegan2.model import ModulatedConv2d
    g = torch.Generator().manual_seed(7)
    f4 = fir.setup_filter([1, 3, 3, 1])
    for tag, up in (('plain', False), ('up2', True)):
        m = ModulatedConv2d(8, 5, 3, 512, upsample=up)
        m.modulation = torch.nn.Identity()
        w = torch.randn(5, 8, 3, 3, generator=g)
        with torch.no_grad():
            m.weight.copy_(w[None])
        x = torch.randn(3, 8, 6, 6, generator=g)
        s = torch.randn(3, 8, generator=g) + 1
        with torch.no_grad():
            yr = m(x, s)
        kw = dict(up=2, resample_filter=f4, flip_weight=False) if up else {}
        # e4e scales W by 1/sqrt(fan_in) before demodulation; demodulation cancels it up to the 1e-8 eps
        yo = conv.modulated_conv2d(x, w, s, padding=1, **kw)
        close(yo, yr, 5e-6, f'modulated_conv2d {tag} vs e4e ModulatedConv2d')
        out[f'e4e_{tag}.x'], out[f'e4e_{tag}.w'], out[f'e4e_{tag}.s'], out[f'e4e_{tag}.y'] = x.numpy(), w.numpy(), s.numpy(), yr.numpy()


def make_small_net(seed=1):
    """64-px net with the same topology as config-f but thin (fast on CPU, small fixtures)."""
    return synthesis.make_generator(64, seed=seed, channel_base=2048, channel_max=512)  # b4 must be 512 wide (utils.py:135)


def pin_driver(R, out):
    print('driver: reference utils.generate_image/get_styles driving the restated network modules')
    G_ref, G_ora = make_small_net(), make_small_net()
    num_ws = G_ref.synthesis.num_ws
    ws = torch.randn(3, num_ws, 512, generator=torch.Generator().manual_seed(2))
    bw_r, bw_o = R.utils.split_ws(G_ref, ws), synthesis.split_ws(G_ora, ws)
    for a, b in zip(bw_r, bw_o):
        close(b, a, 0, 'split_ws')
    S_r, shapes_r = R.utils.get_styles(G_ref, ws, bw_r, torch.device('cpu'))
    S_o, shapes_o = synthesis.get_styles(G_ora, ws, bw_o)
    assert shapes_r == shapes_o, (shapes_r, shapes_o)
    close(S_o, S_r, 1e-6, 'get_styles')
    G3 = make_small_net()
    assert R.utils.get_temp_shapes(G3) == shapes_o == synthesis.get_temp_shapes(make_small_net())
    xs_r, img_r = R.utils.generate_image(G_ref, 100, S_r, shapes_r, 'const', torch.device('cpu'))
    xs_o, img_o = synthesis.generate_image(G_ora, 100, S_r, shapes_o, 'const')
    close(img_o, img_r, 0, 'generate_image img')
    for i, (a, b) in enumerate(zip(xs_r, xs_o)):
        close(b, a, 0, f'generate_image xs[{i}]')
    xs_t, img_t = R.utils.generate_image(G_ref, 2, S_r, shapes_r, 'const', torch.device('cpu'))
    close(synthesis.generate_image(G_ora, 2, S_r, shapes_o, 'const')[1], img_t, 0, 'generate_image until_k=2')
    out['ws'], out['styles'], out['img'] = ws.numpy(), S_r.numpy(), img_r.numpy()
    out['img_k2'] = img_t.numpy()
    out['temp_shapes'] = np.array(shapes_r)
    for i, a in enumerate(xs_r):
        out[f'xs{i}'] = a[:1].numpy()
    print(f'  image stats: mean {img_r.mean():.3f} std {img_r.std():.3f} min {img_r.min():.2f} max {img_r.max():.2f}')
    return G_ref, G_ora, S_r, shapes_r


def pin_clip(R, out, cfg=vit.VIT_B32, seed=0, name='ViT-B/32'):
    print(f'clip: restated {name} vs transformers.CLIPModel (weight copy)')
    model = vit.CLIP(seed=seed, cfg=cfg)
    g = torch.Generator().manual_seed(3)
    images = torch.randn(2, 3, 224, 224, generator=g)
    toks = torch.cat([vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg')])
    with torch.no_grad():
        ei, et = model.encode_image(images), model.encode_text(toks)
    try:
        from transformers import CLIPConfig, CLIPModel
        hcfg = CLIPConfig()  # defaults are ViT-B/32 with quick_gelu
        hcfg.vision_config.patch_size = cfg['vision_patch_size']
        hcfg.vision_config.attn_implementation = hcfg.text_config.attn_implementation = 'eager'
        hf = CLIPModel(hcfg).eval()
        sd = hf.state_dict()
        p = model.p

        def put(k, v):
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            sd[k] = v.clone()
        put('vision_model.embeddings.patch_embedding.weight', p['visual.conv1.weight'])
        put('vision_model.embeddings.class_embedding', p['visual.class_embedding'])
        put('vision_model.embeddings.position_embedding.weight', p['visual.positional_embedding'])
        put('vision_model.pre_layrnorm.weight', p['visual.ln_pre.weight'])
        put('vision_model.pre_layrnorm.bias', p['visual.ln_pre.bias'])
        put('vision_model.post_layernorm.weight', p['visual.ln_post.weight'])
        put('vision_model.post_layernorm.bias', p['visual.ln_post.bias'])
        put('visual_projection.weight', p['visual.proj'].t())
        put('text_model.embeddings.token_embedding.weight', p['token_embedding.weight'])
        put('text_model.embeddings.position_embedding.weight', p['positional_embedding'])
        put('text_model.final_layer_norm.weight', p['ln_final.weight'])
        put('text_model.final_layer_norm.bias', p['ln_final.bias'])
        put('text_projection.weight', p['text_projection'].t())
        for tower, pre, width, layers in (('vision_model', 'visual.transformer', 768, 12), ('text_model', 'transformer', 512, 12)):
            for i in range(layers):
                a, b = f'{tower}.encoder.layers.{i}.', f'{pre}.resblocks.{i}.'
                wq, wk, wv = p[b + 'attn.in_proj_weight'].chunk(3)
                bq, bk, bv = p[b + 'attn.in_proj_bias'].chunk(3)
                for nm, w_, b_ in (('q', wq, bq), ('k', wk, bk), ('v', wv, bv)):
                    put(a + f'self_attn.{nm}_proj.weight', w_)
                    put(a + f'self_attn.{nm}_proj.bias', b_)
                put(a + 'self_attn.out_proj.weight', p[b + 'attn.out_proj.weight'])
                put(a + 'self_attn.out_proj.bias', p[b + 'attn.out_proj.bias'])
                for hfn, on in (('layer_norm1', 'ln_1'), ('layer_norm2', 'ln_2'), ('mlp.fc1', 'mlp.c_fc'), ('mlp.fc2', 'mlp.c_proj')):
                    put(a + hfn + '.weight', p[b + on + '.weight'])
                    put(a + hfn + '.bias', p[b + on + '.bias'])
        hf.load_state_dict(sd)
        with torch.no_grad():
            hi = hf.get_image_features(pixel_values=images)
            ht = hf.get_text_features(input_ids=toks)
        hi = hi if torch.is_tensor(hi) else hi.pooler_output
        ht = ht if torch.is_tensor(ht) else ht.pooler_output
        close(ei, hi, 2e-4, 'encode_image vs transformers')
        close(et, ht, 2e-4, 'encode_text vs transformers')
    except ImportError as e:  # pragma: no cover
        print('  transformers unavailable, cross-check skipped:', e)
    out['tokens'], out['image_features'], out['text_features'] = toks.numpy(), ei.numpy(), et.numpy()  # images = randn(2,3,224,224) seed 3
    return model


def install_stub_clip(R, model, model_b16=None):
    R.clip.load = lambda name, device=None: ({'ViT-B/32': model, 'ViT-B/16': model_b16}[name], None)
    R.clip.tokenize = lambda texts: vit.synthetic_tokens('pos' if 'woman' in texts[0] else 'neg')


POS_TEXT, NEG_TEXT = 'a photo of a face of a feminine woman with no makeup', 'a photo of a face of a masculine man'


def reference_step(R, G, shapes, styles, delta, until_k, l2_reg_coef=0.1):
    """find_direction.py:306-336 executed with the reference's own functions."""
    from torchvision.transforms import CenterCrop, Compose, Resize
    from PIL import Image
    dev = torch.device('cpu')
    mean, std = R.utils.get_mean_std(dev)
    transf = Compose([Resize(224, interpolation=Image.BICUBIC), CenterCrop(224)])
    T = R.fd.S_TRAINABLE_SPACE_CHANNELS
    loss_fn = R.cl.CLIPLoss(dev, POS_TEXT, NEG_TEXT, 'small')
    delta = delta.clone().requires_grad_(True)
    styles_direction = torch.zeros(1, R.fd.N_STYLE_CHANNELS, 512)
    styles_direction[:, T] = delta
    styles2 = styles + styles_direction
    _, img = R.utils.generate_image(G, until_k, styles2, shapes, 'const', dev)
    _, original = R.utils.generate_image(G, until_k, styles, shapes, 'const', dev)
    clip_term = R.fd.compute_clip_loss(img, original, 'default', 'small', 1.0, loss_fn, None, transf, mean, std, dev, POS_TEXT, NEG_TEXT)
    l2 = l2_reg_coef * torch.nn.functional.mse_loss(styles2[:, T], styles[:, T])
    loss = clip_term + l2
    loss.backward()
    return dict(loss=loss.detach(), clip_loss=clip_term.detach(), l2_loss=l2.detach(), grad=delta.grad.clone(),
                img=img.detach(), original_img=original.detach(), unprocessed=R.fd.unprocess(original.detach(), transf, mean, std))


def pin_step(R, model, G_ref, G_ora, S, shapes, out):
    print('step: reference find_direction loop body vs oracle.direction.direction_step (64-px net)')
    install_stub_clip(R, model)
    delta = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(4))
    r = reference_step(R, G_ref, shapes, S, delta, 100)
    loss_fn = direction.CLIPLoss(model, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    o = direction.direction_step(G_ora, shapes, loss_fn, S, delta, 100)
    close(direction.unprocess(r['original_img']), r['unprocessed'], 2e-6, 'unprocess vs find_direction.unprocess+torchvision')
    close(o['img'], r['img'], 0, 'step img')
    close(o['loss'], r['loss'], 1e-6, 'step loss')
    rel = ((o['grad'] - r['grad']).norm() / r['grad'].norm()).item()
    print(f'  grad rel-l2 {rel:.3e}, |grad| {r["grad"].norm():.3e}')
    assert rel < 1e-5
    out['delta'], out['loss'], out['clip_loss'], out['l2_loss'] = delta.numpy(), r['loss'].numpy(), r['clip_loss'].numpy(), r['l2_loss'].numpy()
    out['grad'], out['img'], out['original_img'] = r['grad'].numpy(), r['img'].numpy(), r['original_img'].numpy()
    out['unprocessed'] = r['unprocessed'][:1].numpy()
    out['styles'] = S.numpy()


def pin_step_double(R, model, model_b16, G_ref, G_ora, S, shapes, out):
    """clip_type='double' (the CLI default, find_direction.py:216): the reference's own init_clip_loss / compute_clip_loss
    (find_direction.py:100-122,148-169) on two stub towers vs oracle.direction.DoubleCLIPLoss."""
    print('step (double): reference init_clip_loss + compute_clip_loss with clip_type="double" vs the oracle (64-px net)')
    from torchvision.transforms import CenterCrop, Compose, Resize
    from PIL import Image
    install_stub_clip(R, model, model_b16)
    dev = torch.device('cpu')
    mean, std = R.utils.get_mean_std(dev)
    transf = Compose([Resize(224, interpolation=Image.BICUBIC), CenterCrop(224)])
    T = R.fd.S_TRAINABLE_SPACE_CHANNELS
    l1, l2 = R.fd.init_clip_loss('default', 'double', dev, POS_TEXT, NEG_TEXT)
    assert l1.model is model and l2.model is model_b16
    delta0 = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(4))      # same delta as step64.npz
    delta = delta0.clone().requires_grad_(True)
    styles_direction = torch.zeros(1, R.fd.N_STYLE_CHANNELS, 512)
    styles_direction[:, T] = delta
    styles2 = S + styles_direction
    _, img = R.utils.generate_image(G_ref, 100, styles2, shapes, 'const', dev)
    _, original = R.utils.generate_image(G_ref, 100, S, shapes, 'const', dev)
    clip_term = R.fd.compute_clip_loss(img, original, 'default', 'double', 1.0, l1, l2, transf, mean, std, dev, POS_TEXT, NEG_TEXT)
    reg = 0.1 * torch.nn.functional.mse_loss(styles2[:, T], S[:, T])
    loss = clip_term + reg
    loss.backward()
    pos, neg = vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg')
    loss_fn = direction.DoubleCLIPLoss(direction.CLIPLoss(model, pos, neg), direction.CLIPLoss(model_b16, pos, neg))
    o = direction.direction_step(G_ora, shapes, loss_fn, S, delta0, 100)
    close(o['loss'], loss.detach(), 1e-6, 'double step loss')
    rel = ((o['grad'] - delta.grad).norm() / delta.grad.norm()).item()
    print(f'  grad rel-l2 {rel:.3e}, |grad| {delta.grad.norm():.3e}')
    assert rel < 1e-5
    out['loss'], out['clip_loss'], out['l2_loss'], out['grad'] = (loss.detach().numpy(), clip_term.detach().numpy(), reg.detach().numpy(),
                                                                   delta.grad.numpy())
    out['b16_seed'] = np.array(B16_SEED)     # styles, delta and images are those of step64.npz


B16_SEED = 7


def pin_mapper(R, model, G_ref, G_ora, S, shapes, out):
    """The latent mapper (latent_mappers.py:68-93) and one train_latent_mapper.py:150-190 step: the reference's REAL ``Mapper`` loaded (strict) with
    oracle.mapper.random_mapper_params, its REAL ``find_direction.compute_loss`` (CLIP term through the stub clip module, identity term through the
    hand-built IDLoss of pin_idloss with coefficient 0.6, landmarks coefficient 0, L2 0.1) and autograd down to the mapper's weights, vs
    oracle.mapper on the 64-px network with the styles of step64.npz."""
    print('mapper: reference latent_mappers.Mapper + compute_loss vs oracle.mapper (64-px net)')
    from torchvision.transforms import CenterCrop, Compose, Resize
    from PIL import Image
    op = types.ModuleType('encoder4editing.models.stylegan2.op')          # CUDA-only fused ops: not used by the Mapper (only EqualLinear / PixelNorm names are imported)
    op.FusedLeakyReLU, op.fused_leaky_relu, op.upfirdn2d = torch.nn.Identity, (lambda x, b=None, *a, **k: x), None
    sys.modules.setdefault('encoder4editing.models.stylegan2.op', op)
    sys.path.insert(0, REF)
    from latent_mappers import Mapper
    from id_loss.id_loss import IDLoss
    from id_loss.model_irse import Backbone
    install_stub_clip(R, model)
    dev = torch.device('cpu')
    p = mapper.random_mapper_params(seed=3)
    ref = Mapper(neg_slope=0.01)
    ref.load_state_dict(p, strict=True)
    x = S[:, R.fd.S_TRAINABLE_SPACE_CHANNELS]
    with torch.no_grad():
        close(mapper.mapper_forward(p, x), ref(x), 1e-6, 'Mapper.forward')
    # the reference's objects for compute_loss
    idp = idloss.random_irse50_params(seed=0)
    net = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode='ir_se')
    net.load_state_dict(idp, strict=True)
    net.eval()
    id_ref = IDLoss.__new__(IDLoss)
    torch.nn.Module.__init__(id_ref)
    id_ref.facenet, id_ref.pool, id_ref.face_pool = net, torch.nn.AdaptiveAvgPool2d((256, 256)), torch.nn.AdaptiveAvgPool2d((112, 112))
    mean, std = R.utils.get_mean_std(dev)
    transf = Compose([Resize(224, interpolation=Image.BICUBIC), CenterCrop(224)])
    l1, _ = R.fd.init_clip_loss('default', 'small', dev, POS_TEXT, NEG_TEXT)
    T = R.fd.S_TRAINABLE_SPACE_CHANNELS
    styles = S[:2]
    delta = ref(styles[:, T])                                            # train_latent_mapper.py:155-158
    styles2 = styles.clone()
    styles2[:, T] += delta
    _, img = R.utils.generate_image(G_ref, 100, styles2, shapes, 'const', dev)
    _, original = R.utils.generate_image(G_ref, 100, styles, shapes, 'const', dev)
    loss, ld = R.fd.compute_loss(img, original, transf, mean, std, dev, 'default', 'small', 1.0, l1, None, POS_TEXT, NEG_TEXT, id_ref, 0.6,
                                 None, 0.0, None, 224, styles, styles2, 0.1)
    loss.backward()
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss_fn = direction.CLIPLoss(model, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    o = mapper.mapper_step_loss(G_ora, shapes, loss_fn, pr, styles, 100, id_params=idp, identity_loss_coef=0.6)
    grads = torch.autograd.grad(o['loss'], list(pr.values()))
    close(o['loss'].detach(), loss.detach(), 1e-6, 'mapper step loss')
    close(o['identity_loss'].detach(), ld['identity_loss'].detach(), 1e-6, 'mapper step identity term')
    worst = 0.0
    for (k, _), g in zip(pr.items(), grads):
        gr = dict(ref.named_parameters())[k].grad
        worst = max(worst, ((g - gr).norm() / gr.norm()).item())
    print(f'  loss {loss.item():.6f} (clip {ld["clip_loss"].item():.6f}, identity {ld["identity_loss"].item():.6f}, l2 {ld["l2_loss"].item():.6f}); '
          f'worst parameter-gradient rel-l2 {worst:.2e}')
    assert worst < 1e-4
    out['x'], out['delta'] = x.numpy(), ref(x).detach().numpy()
    out['loss'], out['clip_loss'], out['identity_loss'], out['l2_loss'] = (loss.detach().numpy(), ld['clip_loss'].detach().numpy(),
                                                                           ld['identity_loss'].detach().numpy(), ld['l2_loss'].detach().numpy())
    for k, v in ref.named_parameters():          # full gradients of the biases and of the first / last weight; norms of the other weights (fixture size)
        if k.endswith('bias') or k in ('course_mapping.modulation_module_list.0.fc.weight', 'medium_mapping.modulation_module_list.4.fc.weight'):
            out['grad.' + k] = v.grad.numpy()
        out['gradnorm.' + k] = v.grad.norm().numpy()


def pin_idloss(R, out):
    """The identity loss (find_direction.py:179-180): the reference's REAL id_loss.model_irse.Backbone (IR-SE50, eval mode) loaded (strict)
    with oracle.idloss.random_irse50_params, and the REAL IDLoss.extract_feats / IDLoss.forward run on an instance built without the
    checkpoint file (id_loss/model_ir_se50.pth is not in the tree), vs oracle.idloss; 256-px and 1024-px shaped inputs."""
    print('idloss: reference id_loss.IDLoss / model_irse.Backbone vs oracle.idloss (random IR-SE50 parameters)')
    sys.path.insert(0, REF)
    from id_loss.id_loss import IDLoss
    from id_loss.model_irse import Backbone
    p = idloss.random_irse50_params(seed=0)
    net = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode='ir_se')
    net.load_state_dict(p, strict=True)                      # key names and shapes of the restated parameter dict
    net.eval()
    ref = IDLoss.__new__(IDLoss)                             # IDLoss.__init__ insists on the checkpoint file: build the same object by hand
    torch.nn.Module.__init__(ref)
    ref.facenet, ref.pool, ref.face_pool = net, torch.nn.AdaptiveAvgPool2d((256, 256)), torch.nn.AdaptiveAvgPool2d((112, 112))
    g = torch.Generator().manual_seed(9)
    for tag, res in (('256', 256), ('512', 512)):
        base = torch.nn.functional.interpolate(torch.randn(2, 3, res // 8, res // 8, generator=g), size=res, mode='bicubic', align_corners=False) * 0.6
        y = base.clone()
        y_hat = (base + 0.15 * torch.nn.functional.interpolate(torch.randn(2, 3, res // 16, res // 16, generator=g), size=res, mode='bicubic',
                                                                 align_corners=False)).requires_grad_(True)
        with torch.no_grad():
            fr = ref.extract_feats(y)
        close(idloss.extract_feats(p, y), fr, 1e-5, f'extract_feats {tag} px')
        loss_r, _ = ref(y_hat, y)
        loss_r.backward()
        yh2 = y_hat.detach().clone().requires_grad_(True)
        loss_o = idloss.id_loss(p, yh2, y)
        g_o, = torch.autograd.grad(loss_o, yh2)
        close(loss_o.detach(), loss_r.detach(), 1e-6, f'id loss {tag} px')
        rel = ((g_o - y_hat.grad).norm() / y_hat.grad.norm()).item()
        print(f'  {tag} px: loss {loss_r.item():.6f}, image-gradient rel-l2 {rel:.2e}, |grad| {y_hat.grad.norm():.3e}')
        assert rel < 1e-4
        if res == 256:
            out['y'], out['y_hat'] = y.numpy(), y_hat.detach().numpy()
            out['feats_y'], out['loss'], out['grad'] = fr.numpy(), loss_r.detach().numpy(), y_hat.grad.numpy()
            out['crop112'] = idloss.face_crop(y)[:1].numpy()
        else:
            out['y512_seed'] = np.array(9)
            out['loss512'], out['grad512_norm'] = loss_r.detach().numpy(), y_hat.grad.norm().numpy()
            out['grad512_down'] = torch.nn.functional.avg_pool2d(y_hat.grad, 8).numpy()


def pin_step_nada(R, model, G_ref, G_ora, S, shapes, out):
    """clip_loss_type 'nada' and 'nada_global' (find_direction.py:101-114,150-158): the reference's REAL clip_loss_nada.CLIPLoss
    (clip_loss_nada.py:66-345) driven by its own init_clip_loss / compute_clip_loss, on a stub ``clip`` module (oracle ViT-B/32,
    synthetic tokenizer, the five-transform preprocessing pipeline of clip.load) vs oracle.direction.CLIPLossNADA; 64-px net, the
    styles and delta of step64.npz."""
    print('step (nada): reference clip_loss_nada.CLIPLoss through init_clip_loss / compute_clip_loss vs the oracle (64-px net)')
    from torchvision.transforms import CenterCrop, Compose, Normalize, Resize, ToTensor
    from PIL import Image
    pre = Compose([Resize(224, interpolation=Image.BICUBIC), CenterCrop(224), lambda im: im.convert('RGB'), ToTensor(),
                   Normalize(direction.CLIP_MEAN, direction.CLIP_STD)])
    dummy = types.SimpleNamespace(encode_image=None)
    R.clip.load = lambda name, device=None: ((model if name == 'ViT-B/32' else dummy), pre)
    R.clip.tokenize = lambda texts: vit.synthetic_tokenize(texts)
    dev = torch.device('cpu')
    T = R.fd.S_TRAINABLE_SPACE_CHANNELS
    delta0 = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(4))      # same delta as step64.npz
    for kind in ('nada', 'nada_global'):
        l1, l2 = R.fd.init_clip_loss(kind, 'small', dev, POS_TEXT, NEG_TEXT)
        assert l2 is None and l1.model is model
        delta = delta0.clone().requires_grad_(True)
        styles_direction = torch.zeros(1, R.fd.N_STYLE_CHANNELS, 512)
        styles_direction[:, T] = delta
        styles2 = S + styles_direction
        _, img = R.utils.generate_image(G_ref, 100, styles2, shapes, 'const', dev)
        _, original = R.utils.generate_image(G_ref, 100, S, shapes, 'const', dev)
        clip_term = R.fd.compute_clip_loss(img, original, kind, 'small', 1.0, l1, None, None, None, None, dev, POS_TEXT, NEG_TEXT)
        reg = 0.1 * torch.nn.functional.mse_loss(styles2[:, T], S[:, T])
        loss = clip_term + reg
        loss.backward()
        kw = dict(lambda_direction=0.0, lambda_global=1.0) if kind == 'nada_global' else {}
        o = direction.direction_step(G_ora, shapes, direction.CLIPLossNADA(model, vit.synthetic_tokenize, **kw), S, delta0, 100,
                                     nada_prompts=(NEG_TEXT, POS_TEXT))
        close(o['loss'], loss.detach(), 1e-6, f'{kind} step loss')
        rel = ((o['grad'] - delta.grad).norm() / delta.grad.norm()).item()
        print(f'  {kind}: loss {loss.item():.6f} grad rel-l2 {rel:.3e}, |grad| {delta.grad.norm():.3e}')
        assert rel < 1e-5
        out[kind + '.loss'], out[kind + '.clip_loss'], out[kind + '.grad'] = loss.detach().numpy(), clip_term.detach().numpy(), delta.grad.numpy()
    out['pos_text'], out['neg_text'] = np.array(POS_TEXT), np.array(NEG_TEXT)
    close(direction.nada_preprocess(original), l1.preprocess(original), 2e-6, 'nada_preprocess vs the reference transform pipeline')
    out['preprocessed'] = l1.preprocess(original)[:1].detach().numpy()


def pin_config1(R, model, out):
    print('config 1: FFHQ-256 config-f net, batch 4, one find_direction step through the reference (CPU)')
    G_ref = synthesis.make_generator(256, seed=0)
    G_ora = synthesis.make_generator(256, seed=0)
    ws = torch.randn(129, G_ref.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(0))[:4]
    S, shapes = R.utils.get_styles(G_ref, ws, R.utils.split_ws(G_ref, ws), torch.device('cpu'))
    synthesis.get_temp_shapes(G_ora)
    delta = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(5))
    r = reference_step(R, G_ref, shapes, S, delta, direction.RESOLUTION_TO_K[256])
    loss_fn = direction.CLIPLoss(model, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    o = direction.direction_step(G_ora, shapes, loss_fn, S, delta, direction.RESOLUTION_TO_K[256])
    close(o['img'], r['img'], 0, 'config1 img')
    close(o['loss'], r['loss'], 1e-6, 'config1 loss')
    rel = ((o['grad'] - r['grad']).norm() / r['grad'].norm()).item()
    print(f'  grad rel-l2 {rel:.3e}, |grad| {r["grad"].norm():.3e}; img std {r["img"].std():.3f} '
          f'frac |img|>1: {(r["img"].abs() > 1).float().mean():.4f}')
    assert rel < 1e-5
    out['ws'], out['delta'], out['loss'], out['clip_loss'] = ws.numpy(), delta.numpy(), r['loss'].numpy(), r['clip_loss'].numpy()
    out['grad'] = r['grad'].numpy()
    out['img_crop'] = r['img'][:, :, 96:160, 96:160].numpy()      # 64x64 centre crop of each image
    out['img_mean_std'] = np.array([r['img'].mean().item(), r['img'].std().item()])
    out['img_down'] = torch.nn.functional.avg_pool2d(r['img'], 8).numpy()


CONFIG4_BATCH = 3        # odd on purpose: micro_batch=2 on the GPU side is then a ragged 2 + 1 split


def pin_config4(R, model, out):
    """BASELINE configs[3] (the benchmarked network): FFHQ-1024 config-f net, one find_direction step through the reference's real
    ``utils.generate_image`` + ``compute_clip_loss`` (find_direction.py:306-336), plus the oracle's float64 gradient as the
    reference's own error bar.  Images are stored as 64x64 centre crops and 8x average pools (the full fp32 images are 38 MB)."""
    print(f'config 4: FFHQ-1024 config-f net, batch {CONFIG4_BATCH}, one find_direction step through the reference (CPU)')
    G_ref = synthesis.make_generator(1024, seed=0)
    G_ora = synthesis.make_generator(1024, seed=0)
    ws = torch.randn(CONFIG4_BATCH, G_ref.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(40))
    S, shapes = R.utils.get_styles(G_ref, ws, R.utils.split_ws(G_ref, ws), torch.device('cpu'))
    synthesis.get_temp_shapes(G_ora)
    delta = 0.1 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(41))
    k = direction.RESOLUTION_TO_K[1024]
    r = reference_step(R, G_ref, shapes, S, delta, k)
    loss_fn = direction.CLIPLoss(model, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
    o = direction.direction_step(G_ora, shapes, loss_fn, S, delta, k)
    close(o['img'], r['img'], 0, 'config4 img')
    close(o['loss'], r['loss'], 1e-6, 'config4 loss')
    rel = ((o['grad'] - r['grad']).norm() / r['grad'].norm()).item()
    print(f'  grad rel-l2 {rel:.3e}, |grad| {r["grad"].norm():.3e}; img std {r["img"].std():.3f} '
          f'frac |img|>1: {(r["img"].abs() > 1).float().mean():.4f}')
    assert rel < 1e-5
    # float64 run of the oracle: how far the fp32 reference itself is from the exact gradient
    G64 = synthesis.make_generator(1024, seed=0).double()
    synthesis.get_temp_shapes(G64)
    m64 = vit.CLIP(params=model.p, cfg=model.cfg, dtype=torch.float64)
    if True:
        l64 = direction.CLIPLoss(m64, vit.synthetic_tokens('pos'), vit.synthetic_tokens('neg'))
        o64 = direction.direction_step(G64, shapes, l64, S.double(), delta.double(), k)
        bar = ((r['grad'].double() - o64['grad']).norm() / o64['grad'].norm()).item()
        print(f'  fp32 reference vs float64 oracle: grad rel-l2 {bar:.3e}, loss rel {abs(r["loss"].item() - o64["loss"].item()) / abs(o64["loss"].item()):.3e}')
        out['grad_fp64'], out['loss_fp64'] = o64['grad'].numpy(), o64['loss'].numpy()
    out['ws'], out['delta'], out['loss'], out['clip_loss'] = ws.numpy(), delta.numpy(), r['loss'].numpy(), r['clip_loss'].numpy()
    out['grad'] = r['grad'].numpy()
    for tag, im in (('img', r['img']), ('original', r['original_img'])):
        out[tag + '_crop'] = im[:, :, 480:544, 480:544].numpy()
        out[tag + '_down'] = torch.nn.functional.avg_pool2d(im, 8).numpy()
        out[tag + '_mean_std'] = np.array([im.mean().item(), im.std().item()])
    # generate_fromS.py:147-175,206 on the same network: uint8 canvas original | edited of the first two styles, change_power 2.5,
    # stored as a 16-strided sample plus a 128-row x 256-column full-resolution window around the seam
    styles_direction = torch.zeros(1, R.fd.N_STYLE_CHANNELS, 512)
    styles_direction[:, R.fd.S_TRAINABLE_SPACE_CHANNELS] = delta
    canv = []
    with torch.no_grad():
        for i in range(2):
            halves = []
            for g in (0, 2.5):
                _, img = R.utils.generate_image(G_ref, 100, S[[i]] + styles_direction * g, shapes, 'const', torch.device('cpu'))
                halves.append((img.permute(0, 2, 3, 1) * 127.5 + 128).clamp(0, 255)[0].to(torch.uint8))
            canv.append(torch.cat(halves, dim=1))
    canv = torch.stack(canv)                    # [2, 1024, 2048, 3]
    out['canvas_power'] = np.array(2.5)
    out['canvas_strided'] = canv[:, ::16, ::16].numpy()
    out['canvas_window'] = canv[:, 448:576, 896:1152].numpy()
    out['canvas_sum'] = canv.long().sum(dim=(1, 2)).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--check', action='store_true', help='verify only, do not write fixtures')
    ap.add_argument('--skip-config1', action='store_true')
    ap.add_argument('--skip-config4', action='store_true')
    ap.add_argument('--only-config4', action='store_true', help='(re)write only config4.npz')
    ap.add_argument('--only-mapper', action='store_true', help='(re)write only mapper64.npz')
    ap.add_argument('--only-idloss', action='store_true', help='(re)write only idloss.npz')
    ap.add_argument('--only-nada', action='store_true', help='(re)write only step64_nada.npz')
    ap.add_argument('--only-ops', action='store_true', help='(re)write only ops.npz')
    ap.add_argument('--only-double', action='store_true', help='(re)write only clip_b16.npz and step64_double.npz')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    R = import_reference()
    fx = {k: {} for k in ('ops', 'synth64', 'clip', 'clip_b16', 'step64', 'step64_double', 'step64_nada', 'idloss', 'mapper64', 'config1', 'config4')}
    if args.only_config4:
        model = vit.CLIP(seed=0, cfg=vit.VIT_B32)
        install_stub_clip(R, model)
        pin_config4(R, model, fx['config4'])
        if not args.check:
            np.savez_compressed(os.path.join(GOLD, 'config4.npz'), **fx['config4'])
            print('wrote config4.npz', f'{os.path.getsize(os.path.join(GOLD, "config4.npz")) / 1e6:.2f} MB')
        return
    if args.only_mapper:
        G_ref, G_ora, S, shapes = pin_driver(R, fx['synth64'])
        model = vit.CLIP(seed=0, cfg=vit.VIT_B32)
        pin_mapper(R, model, G_ref, G_ora, S, shapes, fx['mapper64'])
        if not args.check:
            np.savez_compressed(os.path.join(GOLD, 'mapper64.npz'), **fx['mapper64'])
            print('wrote mapper64.npz', f'{os.path.getsize(os.path.join(GOLD, "mapper64.npz")) / 1e6:.2f} MB')
        return
    if args.only_idloss:
        pin_idloss(R, fx['idloss'])
        if not args.check:
            np.savez_compressed(os.path.join(GOLD, 'idloss.npz'), **fx['idloss'])
            print('wrote idloss.npz', f'{os.path.getsize(os.path.join(GOLD, "idloss.npz")) / 1e6:.2f} MB')
        return
    if args.only_nada:
        G_ref, G_ora, S, shapes = pin_driver(R, fx['synth64'])
        model = vit.CLIP(seed=0, cfg=vit.VIT_B32)
        pin_step_nada(R, model, G_ref, G_ora, S, shapes, fx['step64_nada'])
        if not args.check:
            np.savez_compressed(os.path.join(GOLD, 'step64_nada.npz'), **fx['step64_nada'])
            print('wrote step64_nada.npz')
        return
    pin_ops(R, fx['ops'])
    pin_modconv_e4e(R, fx['ops'])
    if args.only_ops:
        if not args.check:
            np.savez_compressed(os.path.join(GOLD, 'ops.npz'), **fx['ops'])
            print('wrote ops.npz')
        return
    G_ref, G_ora, S, shapes = pin_driver(R, fx['synth64'])
    model = pin_clip(R, fx['clip'])
    model_b16 = pin_clip(R, fx['clip_b16'], vit.VIT_B16, B16_SEED, 'ViT-B/16')
    pin_step(R, model, G_ref, G_ora, S, shapes, fx['step64'])
    pin_step_double(R, model, model_b16, G_ref, G_ora, S, shapes, fx['step64_double'])
    pin_step_nada(R, model, G_ref, G_ora, S, shapes, fx['step64_nada'])
    pin_idloss(R, fx['idloss'])
    pin_mapper(R, model, G_ref, G_ora, S, shapes, fx['mapper64'])
    if not args.skip_config1 and not args.only_double:
        pin_config1(R, model, fx['config1'])
    if not args.skip_config4 and not args.only_double:
        pin_config4(R, model, fx['config4'])
    if args.only_double:
        fx = {k: fx[k] for k in ('clip_b16', 'step64_double')}
    if not args.check:
        os.makedirs(GOLD, exist_ok=True)
        for name, d in fx.items():
            if d:
                np.savez_compressed(os.path.join(GOLD, name + '.npz'), **d)
                print('wrote', name + '.npz', f'{os.path.getsize(os.path.join(GOLD, name + ".npz")) / 1e6:.2f} MB')
    print('oracle pinned against reference: OK')


if __name__ == '__main__':
    main()
