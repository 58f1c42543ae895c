"""Forward-only consumer of the hot path: the render loop of ``generate_fromS.py`` (:137-207), batched.

The reference renders one style vector at a time (batch 1), twice (``grad_change`` in {0, change_power}), converts each image to
uint8 on the GPU, copies it to the host and concatenates ``original | edited`` along the width (:174-175, :206).  Here a whole batch of
style vectors goes through the fused synthesis engine in two passes and one small kernel writes both halves of the uint8 canvas; the
mapper / blending / second-generator branches of the reference script are outside the accelerated path.
"""
import torch

from . import _lib, utils


def generate_fromS(G, styles, styles_direction, change_power, device='cuda', noise_mode='const', until_k=100, batch=32, precision='x3p'):
    """styles [M, 26, 512], styles_direction [1, 26, 512] (the ``direction_*.npz`` tensor of find_direction.py:349-351) ->
    uint8 tensor [M, R, 2 R, 3] on ``device``: row i = original image of style i | image of ``styles[i] + change_power * direction``
    (generate_fromS.py:147,166-175,206).  ``out[i].cpu().numpy()`` is what the reference hands to PIL."""
    dev = torch.device(device)
    eng = utils.engine_for(G, dev, precision)
    styles = styles.to(dev, torch.float32)
    direction = styles_direction.to(dev, torch.float32)
    if direction.ndim != 3 or direction.shape[0] != 1 or direction.shape[1:] != styles.shape[1:]:
        raise RuntimeError(f'styles_direction must be [1, {styles.shape[1]}, {styles.shape[2]}], got {tuple(direction.shape)}')
    out = None
    with torch.no_grad(), torch.cuda.device(dev):
        for lo in range(0, styles.shape[0], batch):
            s = styles[lo:lo + batch]
            for j, power in enumerate((0.0, float(change_power))):
                _, img, _ = eng.forward(s + direction * power, until_k=until_k, noise_mode=noise_mode)
                n, _, h, w = img.shape
                if out is None:
                    out = torch.empty([styles.shape[0], h, 2 * w, 3], dtype=torch.uint8, device=dev)
                _lib.call('smc_img_to_uint8', _lib.ptr(img), out[lo:lo + n].data_ptr(), n, h, w, 2 * w, j * w, _lib.stream())
    return out
