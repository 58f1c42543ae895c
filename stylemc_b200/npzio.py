"""(Named npzio, not io: a module called io on sys.path would shadow the standard library when a script of this directory is run directly.)

The reference's on-disk formats either side of the hot path and its optimisation loop.

All files are ``np.savez`` archives with one key:
  ``w``  [M, num_ws, 512] fp32   W+ latents             written by generate_w.py:50-51, read by w_s_converter.py:75
  ``s``  [M, 26, 512] fp32       zero-padded S tensor   written by w_s_converter.py:81-82, read by find_direction.py:260 and
                                                        generate_fromS.py:114
  ``s``  [1, 26, 512] fp32       S-space direction      written by find_direction.py:334 (``direction_last.npz`` checkpoints) and :349-351
                                                        (``direction_<prompt>.npz``), read by generate_fromS.py:125 and ``--resume`` (:266-268)
"""
import math
import os
import warnings

import numpy as np
import torch

N_STYLE_ROWS, STYLE_WIDTH = 26, 512


def _check(a, key, path, lead=None):
    if a.ndim != 3 or a.shape[2] != STYLE_WIDTH or (key == 's' and a.shape[1] != N_STYLE_ROWS) or (lead is not None and a.shape[0] != lead):
        raise RuntimeError(f'{path}: unexpected shape {tuple(a.shape)} for key {key!r}')
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def load_w(path):
    return _check(np.load(path)['w'], 'w', path)


def save_w(path, ws):
    np.savez(path, w=torch.as_tensor(ws).detach().cpu().float().numpy())


def generate_w(G, seeds, truncation_psi=1.0, device='cuda'):
    """generate_w.py:46-51: ``z = RandomState(seed).randn(1, z_dim)`` per seed -> ``G.mapping(z, label, truncation_psi)`` -> W+
    [M, num_ws, 512] (what ``save_w`` writes under the key ``w``).  ``G.mapping`` runs on ``device`` (the lrelu layers use the
    bias_act kernel: CUDA only)."""
    zs = torch.cat([torch.from_numpy(np.random.RandomState(seed).randn(1, G.z_dim)) for seed in seeds])
    G.mapping.to(device)
    with torch.no_grad():
        return G.mapping(zs.to(device, torch.float32), None, truncation_psi=truncation_psi)


def load_styles(path, n=None):
    """[M, 26, 512] fp32 host tensor (generate_fromS.py:114 takes the first ``n`` rows)."""
    s = _check(np.load(path)['s'], 's', path)
    return s if n is None else s[:n]


def save_styles(path, styles):
    np.savez(path, s=torch.as_tensor(styles).detach().cpu().float().numpy())


def direction_path(outdir, text_prompt):
    """find_direction.py:350 / generate_fromS.py:125."""
    return os.path.join(outdir, f'direction_{text_prompt.replace(" ", "_")}.npz')


def load_direction(path):
    return _check(np.load(path)['s'], 's', path, lead=1)


def save_direction(path, styles_direction):
    d = torch.as_tensor(styles_direction).detach().cpu().float()
    if tuple(d.shape) != (1, N_STYLE_ROWS, STYLE_WIDTH):
        raise RuntimeError(f'direction must be [1, {N_STYLE_ROWS}, {STYLE_WIDTH}], got {tuple(d.shape)}')
    np.savez(path, s=d.numpy())


def find_direction(finder, styles_array, batch_size, n_epochs, outdir=None, text_prompt=None, resume=None, seed=0, checkpoint_every=1000,
                   log=None, zero_init='perturb', cache_original=True):
    """The loop of find_direction.py:285-351 around ``DirectionFinder.step``: ``ceil(M / batch) * n_epochs`` iterations, a random
    batch per iteration (:303-304), cosine learning rate (:298-301), ``direction_last.npz`` every ``checkpoint_every`` iterations
    (:333-334), ``direction_<prompt>.npz`` at the end (:349-351); ``resume`` loads a saved direction (:266-270).

    The batch index is drawn from ``np.random.RandomState(seed)``, not the global generator the reference uses: in a data-parallel run
    every rank must draw the same index.  With several ranks each one takes its rows of the batch (``direction.shard_rows``) and only
    rank 0 writes files.  Returns the final direction [1, 26, 512] (host).

    ``zero_init``: what to do when the run would start from delta == 0 (no ``resume``), where the directional loss is 0/0 -- NaN in the
    reference (clip_loss.py:27-28), cos = 0 with a zero gradient here, i.e. a run that never moves: ``'perturb'`` (default) seeds
    delta with ``DirectionFinder.seed_delta`` and warns, ``'raise'`` raises, ``'keep'`` runs as is.

    ``cache_original``: the loop draws the same ``num_batches`` batches over and over, and the CLIP embedding of the un-edited images of a
    batch does not depend on delta: it is computed the first time a batch index is drawn and reused afterwards (``DirectionFinder.step
    (source_key=i)``; identical results, a third less work per later step).  Ignored with the identity term."""
    from . import direction as smc_dir
    n_items = styles_array.shape[0]
    num_batches = math.ceil(n_items / batch_size)
    total = num_batches * n_epochs
    rng = np.random.RandomState(seed)
    rank = torch.distributed.get_rank(finder.group) if finder.world > 1 else 0
    if resume is not None:
        finder.load_direction(load_direction(resume))
    if zero_init not in ('perturb', 'raise', 'keep'):
        raise ValueError("zero_init must be 'perturb', 'raise' or 'keep'")
    if zero_init != 'keep' and total > 0 and not bool(finder.delta.any()):
        msg = ('find_direction starts from delta == 0: the edited and the original image are identical, the directional CLIP loss is '
               '0/0 (NaN in the reference, clip_loss.py:27-28) and its gradient here is 0, so the direction would never move')
        if zero_init == 'raise':
            raise RuntimeError(msg + "; pass resume=... or zero_init='perturb'")
        warnings.warn(msg + '; seeding delta with DirectionFinder.seed_delta()')
        finder.seed_delta()
    if outdir is not None and rank == 0:
        os.makedirs(outdir, exist_ok=True)
    for it in range(1, total + 1):
        lr = smc_dir.cosine_lr(finder.lr, it, total)
        i = rng.randint(0, num_batches)
        batch = styles_array[i * batch_size:(i + 1) * batch_size]
        lo, hi = smc_dir.shard_rows(batch.shape[0], rank, finder.world)
        out = finder.step(batch[lo:hi], lr=lr, global_count=batch.shape[0], source_key=(i, lo, hi) if cache_original else None)
        if outdir is not None and rank == 0 and it % checkpoint_every == checkpoint_every - 1:
            save_direction(os.path.join(outdir, 'direction_last.npz'), finder.direction())
        if log is not None and it % 10 == 0:
            log(it, lr, out)
    final = finder.direction().detach().cpu()
    if outdir is not None and rank == 0:
        save_direction(direction_path(outdir, text_prompt) if text_prompt else os.path.join(outdir, 'direction_last.npz'), final)
    return final


def save_canvases(canvases, outdir, text_prompt, first_index=0):
    """generate_fromS.py:205-206: one ``<prompt>_<i:03d>.jpeg`` (quality 95) per row of the uint8 ``original | edited`` canvas
    [M, R, 2 R, 3] that ``generate.generate_fromS`` returns.  Returns the file paths."""
    from PIL import Image
    arr = torch.as_tensor(canvases).detach().cpu().numpy()
    if arr.dtype != np.uint8 or arr.ndim != 4 or arr.shape[3] != 3:
        raise RuntimeError(f'canvases must be uint8 [M, H, W, 3], got {arr.dtype} {tuple(arr.shape)}')
    os.makedirs(outdir, exist_ok=True)
    paths = []
    for i, a in enumerate(arr, first_index):
        paths.append(os.path.join(outdir, f'{text_prompt.replace(" ", "_")}_{i:03d}.jpeg'))
        Image.fromarray(a, 'RGB').save(paths[-1], quality=95)
    return paths
