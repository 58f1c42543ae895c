import sys, torch
sys.path.insert(0, '.')
from stylemc_b200.ops import upfirdn2d
from stylemc_b200 import _lib
f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for dt in (torch.float32, torch.float16):
    for n, c, w in ((16, 32, 1025), (16, 64, 513), (16, 128, 257), (32, 512, 65)):
        x = torch.randn(n, c, w, w, device='cuda', dtype=dt)
        ms = t(lambda: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4))
        nb = (x.numel() + n * c * (w - 1) * (w - 1)) * x.element_size()
        print(f'{dt} width {w}: {ms:.3f} ms {nb / ms / 1e6:.0f} GB/s {nb / ms / 1e6 / 6551:.2f}', flush=True)
