"""BASELINE configs[4] / BASELINE.md section 3 "second bar": every convolution shape of the config-f synthesis network at 1024 px,
this repo's tcgen05 kernels (split precision x3p = the mode whose gradients match the fp32 reference, and plain fp16 x1) beside the
library path the reference takes on a GPU (conv2d_gradfix.py:35-43 -> F.conv2d / F.conv_transpose2d = cuDNN, fp32 NCHW as the
reference calls it, with TF32 off and on; bf16 channels_last as the library's best case), forward and input gradient.
Plus the op-level modulated_conv2d (forward, and backward to x and the styles) against its eager restatement.

usage (GPU box): python tools/ops_vs_cudnn.py [batch] > profiles/rNN_ops_vs_cudnn.md
Times are CUDA-event averages after warm-up; operands larger than L2 at the high-resolution layers."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemc_b200 import gemm, networks  # noqa: E402
from stylemc_b200.ops import upfirdn2d  # noqa: E402


def timeit(fn, iters=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def library_times(x, w, transposed):
    """cuDNN through PyTorch, as the reference reaches it.  Returns {mode: (fwd ms, dgrad ms)}."""
    out = {}
    for mode in ('fp32', 'tf32', 'bf16_nhwc'):
        torch.backends.cudnn.allow_tf32 = mode == 'tf32'
        torch.backends.cuda.matmul.allow_tf32 = mode == 'tf32'
        xx, ww = x, w
        if mode == 'bf16_nhwc':
            xx = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ww = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        if transposed:
            fwd = lambda: F.conv_transpose2d(xx, ww, stride=2)                     # ww [I, O, 3, 3]
        else:
            fwd = lambda: F.conv2d(xx, ww, padding=1)
        y = fwd()
        gy = torch.randn_like(y)
        xr = xx.detach().requires_grad_(True)
        yr = F.conv_transpose2d(xr, ww, stride=2) if transposed else F.conv2d(xr, ww, padding=1)
        bwd = lambda: torch.autograd.grad(yr, xr, gy, retain_graph=True)
        out[mode] = (timeit(fwd), timeit(bwd))
        del y, gy, xr, yr
    torch.backends.cudnn.allow_tf32 = True
    return out


def ours_conv1(n, c, o, r, prec):
    x3 = prec != 'x1'
    A = torch.randn(2 if x3 else 1, n, r, r, c, device='cuda').half().reshape(-1, r, r, c)
    w = torch.randn(o, c, 3, 3, device='cuda') * 0.05
    Bf, Bb, _, _ = gemm.prepare_weights(w, two=x3, fwd=True, bwd=True)
    d = torch.rand(n, o, device='cuda') + 0.5
    bias = torch.randn(o, device='cuda')
    yh = torch.empty(2 if x3 else 1, n, r, r, o, dtype=torch.float16, device='cuda')
    G = torch.randn(2 if x3 else 1, n, r, r, o, device='cuda').half().reshape(-1, r, r, o)
    gx = torch.empty(2 if x3 else 1, n, r, r, c, dtype=torch.float16, device='cuda')
    acc = 512 if prec == 'x3p' else 0
    p = 'x3' if x3 else 'x1'
    fwd = lambda: gemm.igemm(A, Bf, n, r, r, o, gemm.TAPS_3X3, precision=p, acc_chunk_k=acc, a_plane_stride_imgs=n, b_rows_per_tap=9 * o,
                             row_scale=d, bias=bias, act=1, alpha=0.2, gain=2 ** 0.5, clamp=256.0, out_hi=yh[0], out_lo=yh[1] if x3 else None)
    bwd = lambda: gemm.igemm(G, Bb, n, r, r, c, gemm.TAPS_3X3_DGRAD, precision=p, acc_chunk_k=acc, a_plane_stride_imgs=n, b_rows_per_tap=9 * c,
                             out_hi=gx[0], out_lo=gx[1] if x3 else None)
    return timeit(fwd), timeit(bwd)


def ours_conv0(n, c, o, hin, prec):
    """transposed stride-2 conv as the engine runs it: one problem-group launch of the four parity GEMMs (fp32 planes in x3), and the
    dgrad over the four gradient parity planes."""
    x3 = prec != 'x1'
    A = torch.randn(2 if x3 else 1, n, hin, hin, c, device='cuda').half().reshape(-1, hin, hin, c)
    w = torch.randn(o, c, 3, 3, device='cuda') * 0.05
    Bf, Bb, _, _ = gemm.prepare_weights(w, two=x3, fwd=True, bwd=True)
    d = torch.rand(n, o, device='cuda') + 0.5
    planes = torch.empty([4, n, hin + 1, hin + 1, o], dtype=torch.float32 if x3 else torch.float16, device='cuda')
    taps, problems = [], []
    for q, (r, cc) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        t = gemm.up2_parity_taps(r, cc)
        taps += t
        problems.append((len(t), q * planes[0].numel()))
    acc = 512 if prec == 'x3p' else 0
    p = 'x3' if x3 else 'x1'
    kw = dict(out_f32=planes[0]) if x3 else dict(out_raw=planes[0])
    fwd = lambda: gemm.igemm(A, Bf, n, hin + 1, hin + 1, o, taps, precision=p, acc_chunk_k=acc, a_plane_stride_imgs=n, b_rows_per_tap=9 * o,
                             row_scale=d, problems=problems, **kw)
    gp = torch.randn(2 if x3 else 1, 4 * n, hin + 1, hin + 1, o, device='cuda').half()
    gup = torch.empty([n, hin, hin, c], dtype=torch.float32 if x3 else torch.float16, device='cuda')
    bwd = lambda: gemm.igemm(gp.reshape(-1, hin + 1, hin + 1, o), Bb, n, hin, hin, c, gemm.up2_dgrad_taps(n), precision=p, acc_chunk_k=acc,
                             a_plane_stride_imgs=4 * n, b_rows_per_tap=9 * c, **(dict(out_f32=gup) if x3 else dict(out_raw=gup)))
    return timeit(fwd), timeit(bwd)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    print(f'# Convolution shapes of config-f at 1024 px, batch {n}: stylemc_b200 kernels vs cuDNN through PyTorch (B200)\n')
    print('`python tools/ops_vs_cudnn.py` -- ms per call (CUDA events).  ours x3p = split fp16 operands + promoted accumulation (fp32-grade, the '
          'benchmarked mode; includes the fused demodulation / bias / lrelu / clamp epilogue and the hi+lo plane stores); ours x1 = plain fp16 '
          'operands.  cuDNN: fp32 NCHW as the reference calls it (TF32 off / on), bf16 channels_last = library best case (no epilogue).  '
          'conv0 = transposed stride-2 3x3 (ours: the four parity GEMMs in one launch; FIR not included on either side).\n')
    print('| layer | Cin -> Cout @ out res | GFLOP | ours x3p fwd | ours x1 fwd | cuDNN fp32 | cuDNN tf32 | cuDNN bf16 nhwc | x3p / tf32 | ours x3p dgrad | ours x1 dgrad | '
          'cuDNN fp32 dgrad | cuDNN tf32 dgrad | cuDNN bf16 dgrad | x3p / tf32 |')
    print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    ch = lambda r: min(32768 // r, 512)
    for r in (8, 16, 32, 64, 128, 256, 512, 1024):
        for name in ('conv0', 'conv1'):
            cin = ch(r // 2) if name == 'conv0' else ch(r)
            cout = ch(r)
            hin = r // 2 if name == 'conv0' else r
            x = torch.randn(n, cin, hin, hin, device='cuda')
            w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.05
            lib = library_times(x, w.transpose(0, 1).contiguous() if name == 'conv0' else w, name == 'conv0')
            del x, w
            torch.cuda.empty_cache()
            f = ours_conv0 if name == 'conv0' else ours_conv1
            o3 = f(n, cin, cout, hin, 'x3p')
            torch.cuda.empty_cache()
            o1 = f(n, cin, cout, hin, 'x1')
            torch.cuda.empty_cache()
            gf = 2 * 9 * cin * cout * hin * hin * n / 1e9
            print(f'| b{r}.{name} | {cin} -> {cout} @ {r} | {gf:.1f} | {o3[0]:.3f} | {o1[0]:.3f} | {lib["fp32"][0]:.3f} | {lib["tf32"][0]:.3f} | '
                  f'{lib["bf16_nhwc"][0]:.3f} | {o3[0] / lib["tf32"][0]:.2f} | {o3[1]:.3f} | {o1[1]:.3f} | {lib["fp32"][1]:.3f} | {lib["tf32"][1]:.3f} | '
                  f'{lib["bf16_nhwc"][1]:.3f} | {o3[1] / lib["tf32"][1]:.2f} |', flush=True)

    print('\n## op-level modulated_conv2d (networks.modulated_conv2d, NCHW fp32 in / out) vs its eager restatement on cuDNN (TF32 on)\n')
    print('| shape | ours fwd | eager fwd | ours fwd+bwd (x, styles) | eager fwd+bwd |\n|---|---|---|---|---|')
    f4 = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')

    def eager(x, w, s, up):
        q = w.square().sum(dim=[2, 3])
        dco = (s.square() @ q.t() + 1e-8).rsqrt()
        xm = x * s[:, :, None, None]
        if up:
            y = F.conv_transpose2d(xm, w.transpose(0, 1), stride=2)
            y = F.conv2d(F.pad(y, [1, 1, 1, 1]), (f4 * 4)[None, None].repeat(y.shape[1], 1, 1, 1), groups=y.shape[1])
        else:
            y = F.conv2d(xm, w, padding=1)
        return y * dco[:, :, None, None]

    for (nn, c, o, r, up) in ((8, 512, 512, 64, False), (8, 128, 128, 256, False), (8, 64, 32, 512, True), (8, 32, 32, 1024, False)):
        x = torch.randn(nn, c, r, r, device='cuda', requires_grad=True)
        w = torch.randn(o, c, 3, 3, device='cuda') * 0.05
        s = (torch.randn(nn, c, device='cuda') + 1).requires_grad_(True)
        ours = lambda: networks.modulated_conv2d(x, w, s, padding=1, up=2 if up else 1, resample_filter=f4 if up else None, flip_weight=not up)
        ref = lambda: eager(x, w, s, up)

        def fb(fn):
            y = fn()
            torch.autograd.grad(y, (x, s), torch.ones_like(y))
        with torch.no_grad():
            t_of, t_ef = timeit(ours, 3), timeit(ref, 3)
        t_ob, t_eb = timeit(lambda: fb(ours), 3), timeit(lambda: fb(ref), 3)
        print(f'| n{nn} {c}->{o} @{r}{" up2" if up else ""} | {t_of:.3f} | {t_ef:.3f} | {t_ob:.3f} | {t_eb:.3f} |', flush=True)
        del x, w, s
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
