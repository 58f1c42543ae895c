"""Op microbench (BASELINE configs[4]): the two reference plugins' replacements, `bias_act` and `upfirdn2d`, through the op-level API
(stylemc_b200.ops), at the config-f resolutions / channel counts.  Prints achieved GB/s = algorithmic bytes (read x [+ aux] + write y)
/ CUDA-event time, and the fraction of the measured HBM peak (MEASURED_PEAKS.json).  Inputs are larger than L2.
usage: python tools/op_bench.py [> profiles/rNN_ops.md]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemc_b200.ops import bias_act, upfirdn2d  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    SRC = 'measured'
except Exception:
    HBM, SRC = 6650.0, 'fallback'


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def row(name, shape, dtype, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(f'| {name} | {tuple(shape)} {str(dtype).replace("torch.", "")} | {nbytes / 1e6:.0f} | {ms:.3f} | {gbs:.0f} | {gbs / HBM:.2f} |', flush=True)


def main():
    print(f'# Op microbench on B200: achieved HBM bandwidth of the plugin replacements (peak {HBM:.0f} GB/s, {SRC})\n')
    print('| op | input | algorithmic MB | ms | GB/s | fraction of peak |\n|---|---|---|---|---|---|')
    f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
    for res, c, n in [(256, 128, 16), (512, 64, 16), (1024, 32, 16), (64, 512, 32)]:
        for dtype in (torch.float32, torch.float16):
            es = 2 if dtype == torch.float16 else 4
            x = torch.randn(n, c, res, res, device='cuda', dtype=dtype)
            b = torch.randn(c, device='cuda', dtype=dtype)
            ms = timeit(lambda: bias_act.bias_act(x, b, act='lrelu', clamp=256))
            row('bias_act lrelu fwd (gain sqrt2, clamp 256)', x.shape, dtype, 2 * x.numel() * es, ms)
            ms = timeit(lambda: bias_act.bias_act(x, b, act='linear', clamp=256))
            row('bias_act linear fwd (clamp 256)', x.shape, dtype, 2 * x.numel() * es, ms)
            xr = x.detach().clone().requires_grad_(True)
            y = bias_act.bias_act(xr, b, act='lrelu', clamp=256)
            g = torch.randn_like(y)
            ms = timeit(lambda: torch.autograd.grad(y, xr, g, retain_graph=True))
            row('bias_act lrelu bwd (dy, y -> dx)', x.shape, dtype, 3 * x.numel() * es, ms)
            del xr, y, g
            # the FIR after the transposed conv: 4x4, up = down = 1, pad 1, gain 4, on (2H+1)^2
            t = torch.randn(n, c, res + 1, res + 1, device='cuda', dtype=dtype)
            ms = timeit(lambda: upfirdn2d.upfirdn2d(t, f, padding=[1, 1, 1, 1], gain=4))
            row('upfirdn2d 4x4 up1 pad1 (conv0 FIR)', t.shape, dtype, (t.numel() + n * c * res * res) * es, ms)
            del t, x
            torch.cuda.empty_cache()
        # the skip-image path: C = 3, up = 2 and its transpose (down = 2)
        img = torch.randn(n * 4, 3, res // 2, res // 2, device='cuda')
        ms = timeit(lambda: upfirdn2d.upsample2d(img, f))
        row('upfirdn2d.upsample2d (skip image, up 2)', img.shape, torch.float32, (img.numel() + img.numel() * 4) * 4, ms)
        big = torch.randn(n * 4, 3, res, res, device='cuda')
        ms = timeit(lambda: upfirdn2d.upfirdn2d(big, f, down=2, padding=[1, 1, 1, 1], flip_filter=True, gain=4))
        row('upfirdn2d down 2 (transpose of upsample2d)', big.shape, torch.float32, (big.numel() + big.numel() // 4) * 4, ms)
        del img, big


if __name__ == '__main__':
    main()
