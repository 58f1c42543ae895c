"""GPU probe of the CTA-pair (cta_group::2) launch of csrc/hconv.cu: correctness against the single-CTA kernel and fp64, then timing of both
on the BN = 128 layer shapes.  Run each stage under `timeout` (a protocol bug in a first bring-up shows up as a hang, not as a wrong number)."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from stylemc_b200 import _lib, gemm  # noqa: E402


def case(n, c, o, h, w, x3, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(n, c, h, w, device='cuda', generator=g)
    wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
    if not x3:
        x, wt = x.half().float(), wt.half().float()
    A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), x3).reshape(-1, h, w, c)
    B, _, _, _ = gemm.prepare_weights(wt, two=x3)
    out = torch.empty(n, h, w, o, device='cuda')

    def run():
        gemm.igemm(A, B, n, h, w, o, gemm.TAPS_3X3, precision='x3' if x3 else 'x1', acc_chunk_k=512 if x3 else 0, a_plane_stride_imgs=n,
                   b_rows_per_tap=9 * o, out_f32=out)
        return out
    return x, wt, run


def timeit(fn, iters=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


what = sys.argv[1] if len(sys.argv) > 1 else 'check'
_lib.call('smc_igemm_config', 0, 2)
if what == 'check':
    for (n, c, o, h, w) in [(2, 64, 128, 16, 16), (2, 128, 128, 40, 40), (4, 256, 256, 32, 32), (6, 128, 128, 64, 64)]:
        for x3 in (False, True):
            x, wt, run = case(n, c, o, h, w, x3)
            _lib.call('smc_igemm_config', 7, 0)
            a = run().clone()
            _lib.call('smc_igemm_config', 7, 1)
            b = run().clone()
            torch.cuda.synchronize()
            ref = F.conv2d(x.double(), wt.double(), padding=1).permute(0, 2, 3, 1)
            err = ((b.double() - ref).abs().max() / ref.abs().max()).item()
            print(f'n{n} c{c} o{o} {h}x{w} x3={x3}: pair == single: {torch.equal(a, b)}  max diff {(a - b).abs().max().item():.2e}  rel err vs fp64 {err:.2e}', flush=True)
else:
    n = 16
    for (c, o, r) in [(512, 512, 32), (512, 512, 64), (256, 256, 128), (128, 128, 256), (256, 128, 256)]:
        x, wt, run = case(n, c, o, r, r, True)
        fl = 2.0 * 9 * c * o * r * r * n
        row = {}
        for pair in (0, 1, 0, 1):
            _lib.call('smc_igemm_config', 7, pair)
            row[f'pair{pair}'] = row.get(f'pair{pair}', []) + [timeit(run)]
        print(f'c{c} o{o} {r}x{r} n{n} x3p:', {k: [f'{v:.3f} ms ({fl / v / 1e9:.0f} TF/s alg)' for v in vs] for k, vs in row.items()}, flush=True)
        del x, wt, run
        torch.cuda.empty_cache()
