"""S-space driver: drop-in for the reference's ``utils.py`` functions on the hot path.

``generate_image(G, until_k, styles, temp_shapes, noise_mode, device)`` (utils.py:161-216) keeps its signature and return
value ``(xs, img)`` but runs the fused engine of ``stylemc_b200.synthesis`` instead of calling the network modules layer by
layer.  ``G`` is any object with the attribute structure of the unpickled generator (``G.synthesis.b{res}.conv0/conv1/torgb``
with ``weight``, ``bias``, ``noise_const``, ``noise_strength``; names per legacy.py:173-202); its frozen parameters are
converted once and cached.  ``split_ws`` / ``get_temp_shapes`` / ``get_styles`` (utils.py:77-158) are host logic kept for the
S on-disk format.
"""
import weakref

import torch

from . import synthesis

_engines = weakref.WeakKeyDictionary()


def engine_for(G, device='cuda', precision='x3p'):
    """The (cached) execution plan of a frozen generator.  Parameters are snapshotted on first use."""
    per_g = _engines.setdefault(G, {})
    key = (str(torch.device(device)), precision)
    if key not in per_g:
        per_g[key] = synthesis.SynthesisEngine(G, device=device, precision=precision)
    return per_g[key]


def split_ws(G, ws):
    """utils.py:77-87."""
    block_ws, w_idx = [], 0
    ws = ws.to(torch.float32)
    for res in G.synthesis.block_resolutions:
        block = getattr(G.synthesis, f'b{res}')
        block_ws.append(ws.narrow(1, w_idx, block.num_conv + block.num_torgb))
        w_idx += block.num_conv
    return block_ws


def _layers(block):
    return [block.conv1, block.torgb] if block.in_channels == 0 else [block.conv0, block.conv1, block.torgb]


def get_temp_shapes(G):
    """utils.py:100-120: per block (C_conv0, C_conv1, C_torgb); every ``affine`` becomes Identity (S is fed directly)."""
    shapes = []
    for res in G.synthesis.block_resolutions:
        layers = _layers(getattr(G.synthesis, f'b{res}'))
        c = [layer.affine.weight.shape[0] for layer in layers]
        shapes.append((c[0], c[0], c[1]) if len(c) == 2 else tuple(c))
        for layer in layers:
            layer.affine = torch.nn.Identity()
    return shapes


def get_styles(G, ws, block_ws, device):
    """utils.py:123-158: S [M, 26, 512] zero padded, row j = affine_j(w).  Tiny FCs run once per seed set."""
    styles = torch.zeros(ws.shape[0], synthesis.N_STYLE_ROWS, synthesis.STYLE_WIDTH, device=device)
    shapes, row = [], 0
    with torch.no_grad():
        for res, cur in zip(G.synthesis.block_resolutions, block_ws):
            layers = _layers(getattr(G.synthesis, f'b{res}'))
            c = [layer.affine.weight.shape[0] for layer in layers]
            shapes.append((c[0], c[0], c[1]) if len(c) == 2 else tuple(c))
            for j, layer in enumerate(layers):
                styles[:, row, :c[j]] = layer.affine(cur[:, j, :].to(device))
                layer.affine = torch.nn.Identity()
                row += 1
    return styles, shapes


def generate_image(G, until_k, styles, temp_shapes, noise_mode, device, use_blending=False, xs_original=None, masks_dict=None,
                   precision='x3p'):
    """utils.py:161-216.  Returns (xs, img): per-block feature maps [N, C, res, res] fp32 and the running skip image.

    ``temp_shapes`` is accepted for signature compatibility and checked against the network."""
    if use_blending:
        raise RuntimeError('feature blending (utils.py:189-205) is outside the accelerated path')
    eng = engine_for(G, device, precision)
    if temp_shapes is not None:
        for k, blk in enumerate(eng.blocks):
            want = (blk.conv1.cin if blk.conv0 is None else blk.conv0.cin, blk.conv1.cin, blk.torgb.cin)
            if tuple(temp_shapes[k]) != want:
                raise RuntimeError(f'temp_shapes[{k}] = {tuple(temp_shapes[k])} does not match the network {want}')
    xs, img, _ = eng.forward(styles.to(device), until_k=until_k, noise_mode=noise_mode, want_xs=True)
    return xs, img


def style_layer_names(G):
    """Layer name of every used row of the S tensor, in row order (utils.py:133-155 walks the blocks the same way):
    ``['b4.conv1', 'b4.torgb', 'b8.conv0', 'b8.conv1', 'b8.torgb', ...]`` (26 names for the 1024-px network, 20 for 256 px)."""
    names = []
    for res in G.synthesis.block_resolutions:
        block = getattr(G.synthesis, f'b{res}')
        names += [f'b{res}.{n}' for n in (('conv1', 'torgb') if block.in_channels == 0 else ('conv0', 'conv1', 'torgb'))]
    return names


def styles_dict(G, styles, temp_shapes):
    """Per-layer view of the S tensor: ``{'b4.conv1': styles[:, 0, :512], ...}`` (views, no copies).  The reference's container stays the
    zero-padded ``[N, 26, 512]`` tensor plus ``temp_shapes`` (utils.py:123-158); this is the "styles dict" reading of the same data."""
    widths = [c for shape, res in zip(temp_shapes, G.synthesis.block_resolutions)
              for c in (shape[1:] if getattr(G.synthesis, f'b{res}').in_channels == 0 else shape)]
    names = style_layer_names(G)
    if len(widths) != len(names) or len(names) > styles.shape[1]:
        raise RuntimeError('temp_shapes does not match the network')
    return {name: styles[:, row, :c] for row, (name, c) in enumerate(zip(names, widths))}


def styles_from_dict(G, per_layer, temp_shapes=None):
    """Inverse of ``styles_dict``: the zero-padded ``[N, 26, 512]`` tensor from per-layer style vectors."""
    names = style_layer_names(G)
    if set(per_layer) != set(names):
        raise RuntimeError(f'expected styles for {names}, got {sorted(per_layer)}')
    first = per_layer[names[0]]
    styles = torch.zeros(first.shape[0], synthesis.N_STYLE_ROWS, synthesis.STYLE_WIDTH, dtype=torch.float32, device=first.device)
    for row, name in enumerate(names):
        s = per_layer[name]
        styles[:, row, :s.shape[1]] = s
    if temp_shapes is not None:
        for name, view in styles_dict(G, styles, temp_shapes).items():
            if view.shape[1] != per_layer[name].shape[1]:
                raise RuntimeError(f'{name}: {per_layer[name].shape[1]} style channels, the network has {view.shape[1]}')
    return styles
