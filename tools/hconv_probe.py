"""GPU probe for the halo-tile conv kernel (hconv.cu): correctness against torch (fp64) for both UMMA base_offset modes,
then timing of old (igemm.cu) vs new path on the BASELINE layer shapes.  Usage: python tools/hconv_probe.py [check|time] ..."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from stylemc_b200 import _lib, gemm  # noqa: E402


def cfg(key, val):
    _lib.call('smc_igemm_config', key, val)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def conv_case(n, c, o, h, w, x3, acc_k=0, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(n, c, h, w, device='cuda', generator=g)
    wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
    if not x3:
        x, wt = x.half().float(), wt.half().float()
    A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), x3).reshape(-1, h, w, c)
    Bm = wt.permute(2, 3, 0, 1).reshape(9 * o, c)
    B = gemm.split_planes(Bm, x3).reshape(-1, c)
    out = torch.empty(n, h, w, o, device='cuda')

    def run():
        gemm.igemm(A, B, n, h, w, o, gemm.TAPS_3X3, precision='x3' if x3 else 'x1', acc_chunk_k=acc_k,
                   a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=out)
        return out
    return x, wt, run


def check():
    ok = True
    for (n, c, o, h, w) in [(2, 64, 64, 32, 32), (1, 128, 128, 48, 40), (2, 64, 128, 64, 64), (1, 192, 64, 33, 70)]:
        for x3 in (False, True):
            x, wt, run = conv_case(n, c, o, h, w, x3, acc_k=512 if x3 else 0)
            ref = F.conv2d(x.double(), wt.double(), padding=1).permute(0, 2, 3, 1)
            res = {}
            cfg(0, 0)
            res['old'] = rel(run(), ref)
            for bo in (1, 0):
                cfg(0, 2); cfg(1, bo)
                res[f'new_bo{bo}'] = rel(run(), ref)
            cfg(1, 1)
            print(f'conv n{n} c{c} o{o} {h}x{w} x3={x3}:', {k: f'{v:.2e}' for k, v in res.items()}, flush=True)
            ok &= res['new_bo1'] < 1e-4
    print('CHECK', 'PASS' if ok else 'FAIL')


def bench_one(run, iters=5):
    run(); run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timing():
    n = 16
    for (c, o, r) in [(512, 512, 32), (512, 512, 64), (256, 256, 128), (128, 128, 256), (64, 64, 512), (32, 32, 1024), (64, 32, 512)]:
        x, wt, run = conv_case(n, c, o, r, r, True, acc_k=512)
        flops = 2.0 * 9 * c * o * r * r * n
        row = {}
        cfg(0, 0)
        row['old'] = bench_one(run)
        for (wt_, nb) in [(32, 4), (64, 3), (62, 4), (30, 4), (126, 4)]:
            cfg(0, 2); cfg(3, wt_); cfg(2, nb)
            try:
                row[f'new_wt{wt_}_nb{nb}'] = bench_one(run)
            except RuntimeError as ex:
                row[f'new_wt{wt_}_nb{nb}'] = float('nan')
        cfg(3, 32); cfg(2, 4)
        print(f'c{c} o{o} {r}x{r} n{n} x3p:', {k: f'{v:.3f} ms {flops / v / 1e9:.0f} TF/s(alg)' for k, v in row.items()}, flush=True)
        del x, wt, run
        torch.cuda.empty_cache()


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'check'
    if what in ('check', 'all'):
        check()
    if what in ('time', 'all'):
        timing()


def drain_sensitivity():
    n = 16
    for (c, o, r) in [(512, 512, 64), (128, 128, 256)]:
        for (x3, acc_k) in [(True, 512), (True, 1152), (True, 2304), (True, 0), (False, 0)]:
            x, wt, run = conv_case(n, c, o, r, r, x3, acc_k=acc_k)
            cfg(0, 2)
            ms = bench_one(run)
            ref = F.conv2d(x[:2].double(), wt.double(), padding=1).permute(0, 2, 3, 1)
            err = rel(run()[:2], ref)
            flops = 2.0 * 9 * c * o * r * r * n
            print(f'c{c} {r}x{r} x3={x3} acc_k={acc_k}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TF/s(alg)  err {err:.2e}', flush=True)
            del x, wt, run
            torch.cuda.empty_cache()


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'drain':
    drain_sensitivity()


def rms_err():
    """RMS (not max) error of one x3p conv, both kernels, vs fp64."""
    for (n, c, o, r) in [(4, 64, 64, 32), (4, 32, 32, 64), (2, 512, 512, 32), (2, 128, 128, 64)]:
        for acc_k in (512, 0):
            x, wt, run = conv_case(n, c, o, r, r, True, acc_k=acc_k, seed=3)
            ref = F.conv2d(x.double(), wt.double(), padding=1).permute(0, 2, 3, 1)
            row = {}
            for mode in (0, 2):
                cfg(0, mode)
                out = run().double()
                d = out - ref
                row[mode] = (d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(), (d.mean() / ref.abs().mean()).item(), (d * ref.sign()).mean().item() / ref.abs().mean().item()
            print(f'c{c} o{o} {r}x{r} acc_k={acc_k}: ' + '  '.join(f'mode{m}: rms {v[0]:.2e} bias {v[1]:+.1e} shrink {v[2]:+.1e}' for m, v in row.items()), flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'rms':
    rms_err()


def ab_compare():
    """old (igemm.cu) vs new (hconv.cu) vs fp64 on x3p convs: forward taps and dgrad taps, gaussian and heavy-tailed inputs."""
    for (n, c, o, r) in [(4, 64, 64, 32), (4, 32, 32, 64)]:
        for kind in ('fwd', 'dgrad'):
            for dist in ('gauss', 'heavy'):
                g = torch.Generator(device='cuda').manual_seed(5)
                x = torch.randn(n, c, r, r, device='cuda', generator=g)
                if dist == 'heavy':
                    x = x * torch.exp(2.0 * torch.randn(n, c, r, r, device='cuda', generator=g)) * (torch.rand(n, c, r, r, device='cuda', generator=g) > 0.5)
                wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
                A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(-1, r, r, c)
                if kind == 'fwd':
                    Bm = wt.permute(2, 3, 0, 1).reshape(9 * o, c)
                    taps, nout = gemm.TAPS_3X3, o
                    ref = F.conv2d(x.double(), wt.double(), padding=1).permute(0, 2, 3, 1)
                else:
                    wt = torch.randn(c, o, 3, 3, device='cuda', generator=g) * 0.05      # [O=c, I=o]: dgrad maps c -> o channels
                    Bm = wt.permute(2, 3, 1, 0).reshape(9 * o, c)
                    taps, nout = gemm.TAPS_3X3_DGRAD, o
                    xr = torch.zeros(n, o, r, r, device='cuda', dtype=torch.float64, requires_grad=True)
                    F.conv2d(xr, wt.double(), padding=1).backward(x.double())
                    ref = xr.grad.permute(0, 2, 3, 1)
                B = gemm.split_planes(Bm.contiguous(), True).reshape(-1, c)
                outs = {}
                for mode in (0, 2):
                    cfg(0, mode)
                    out = torch.empty(n, r, r, nout, device='cuda')
                    gemm.igemm(A, B, n, r, r, nout, taps, precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n, b_rows_per_tap=9 * nout, out_f32=out)
                    outs[mode] = out.double()
                rms = ref.pow(2).mean().sqrt()
                e0 = ((outs[0] - ref).pow(2).mean().sqrt() / rms).item()
                e2 = ((outs[2] - ref).pow(2).mean().sqrt() / rms).item()
                d = (outs[2] - outs[0]).abs()
                print(f'c{c} o{o} {r}x{r} {kind} {dist}: old rms {e0:.2e} new rms {e2:.2e}  new-old max {d.max().item() / rms.item():.2e} (at {tuple(int(v) for v in torch.nonzero(d == d.max())[0])})', flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'ab':
    ab_compare()


def ab_epilogue():
    import math
    for (n, c, o, r) in [(3, 64, 64, 32), (3, 32, 32, 64), (3, 128, 128, 32)]:
        g = torch.Generator(device='cuda').manual_seed(5)
        x = torch.randn(n, c, r, r, device='cuda', generator=g)
        wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
        d = torch.rand(n, o, device='cuda', generator=g) + 0.5
        bias = torch.randn(o, device='cuda', generator=g) * 0.1
        noise = torch.randn(r, r, device='cuda', generator=g) * 0.1
        A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(-1, r, r, c)
        B = gemm.split_planes(wt.permute(2, 3, 0, 1).reshape(9 * o, c).contiguous(), True).reshape(-1, c)
        u = F.conv2d(x.double(), wt.double(), padding=1)
        z = u * d.double()[:, :, None, None] + noise.double() + bias.double()[None, :, None, None]
        ref = (F.leaky_relu(z, 0.2) * math.sqrt(2)).clamp(-256, 256).permute(0, 2, 3, 1)
        outs = {}
        for mode in (0, 2):
            cfg(0, mode)
            y = torch.zeros(2, n, r, r, o, device='cuda', dtype=torch.float16)
            gemm.igemm(A, B, n, r, r, o, gemm.TAPS_3X3, precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n, b_rows_per_tap=9 * o,
                       row_scale=d, bias=bias, noise=noise, noise_strides=(r, 1), act=1, alpha=0.2, gain=math.sqrt(2), clamp=256.0,
                       out_hi=y[0], out_lo=y[1])
            outs[mode] = y[0].double() + y[1].double()
        rms = ref.pow(2).mean().sqrt()
        for mode in (0, 2):
            dd = (outs[mode] - ref).abs()
            print(f'c{c} o{o} {r}x{r} mode{mode}: rms err {(dd.pow(2).mean().sqrt() / rms).item():.2e} max {dd.max().item() / rms.item():.2e} at {tuple(int(v) for v in torch.nonzero(dd == dd.max())[0])}'
                  f'  per-channel-half rms: {[f"{(dd[..., k * (o // 4):(k + 1) * (o // 4)].pow(2).mean().sqrt() / rms).item():.1e}" for k in range(4)]}', flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'abe':
    ab_epilogue()


def one(c, o, r, n=16):
    x, wt, run = conv_case(n, c, o, r, r, True, acc_k=512)
    cfg(0, 2)
    for _ in range(3):
        run()
    torch.cuda.synchronize()


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'one':
    one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))


def small_res():
    """x3p accuracy of the per-tap kernel at tiny spatial sizes (several images per 128-row tile)."""
    for (n, c, o, r) in [(4, 512, 512, 4), (4, 512, 512, 8), (4, 512, 512, 16), (4, 512, 512, 32)]:
        for kind in ('fwd', 'dgrad', 'up2dgrad', 'up2fwd'):
            g = torch.Generator(device='cuda').manual_seed(5)
            wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
            if kind in ('fwd', 'dgrad'):
                x = torch.randn(n, c if kind == 'fwd' else o, r, r, device='cuda', generator=g)
                A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(-1, r, r, x.shape[1])
                if kind == 'fwd':
                    B = gemm.split_planes(wt.permute(2, 3, 0, 1).reshape(9 * o, c).contiguous(), True).reshape(-1, c)
                    ref = F.conv2d(x.double(), wt.double(), padding=1).permute(0, 2, 3, 1)
                    out = torch.empty(n, r, r, o, device='cuda')
                    gemm.igemm(A, B, n, r, r, o, gemm.TAPS_3X3, precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n, b_rows_per_tap=9 * o, out_f32=out)
                else:
                    B = gemm.split_planes(wt.permute(2, 3, 1, 0).reshape(9 * c, o).contiguous(), True).reshape(-1, o)
                    xr = torch.zeros(n, c, r, r, device='cuda', dtype=torch.float64, requires_grad=True)
                    F.conv2d(xr, wt.double(), padding=1).backward(x.double())
                    ref = xr.grad.permute(0, 2, 3, 1)
                    out = torch.empty(n, r, r, c, device='cuda')
                    gemm.igemm(A, B, n, r, r, c, gemm.TAPS_3X3_DGRAD, precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n, b_rows_per_tap=9 * c, out_f32=out)
            elif kind == 'up2dgrad':
                h = r
                gy = torch.randn(n, o, 2 * h + 1, 2 * h + 1, device='cuda', generator=g)
                gyl = F.pad(gy, (0, 1, 0, 1))
                gp32 = torch.stack([gyl[:, :, rr::2, cc::2] for rr in (0, 1) for cc in (0, 1)])
                gp = gemm.split_planes(gp32.permute(0, 1, 3, 4, 2).contiguous(), True)
                B = gemm.split_planes(wt.permute(2, 3, 1, 0).reshape(9 * c, o).contiguous(), True).reshape(-1, o)
                out = torch.empty(n, h, h, c, device='cuda')
                gemm.igemm(gp.reshape(-1, h + 1, h + 1, o), B, n, h, h, c, gemm.up2_dgrad_taps(n), precision='x3', acc_chunk_k=512,
                           a_plane_stride_imgs=4 * n, b_rows_per_tap=9 * c, out_f32=out)
                xr = torch.zeros(n, c, h, h, device='cuda', dtype=torch.float64, requires_grad=True)
                F.conv_transpose2d(xr, wt.double().transpose(0, 1), stride=2).backward(gy.double())
                ref = xr.grad.permute(0, 2, 3, 1)
            else:
                h = r
                x = torch.randn(n, c, h, h, device='cuda', generator=g)
                A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(-1, h, h, c)
                B = gemm.split_planes(wt.permute(2, 3, 0, 1).reshape(9 * o, c).contiguous(), True).reshape(-1, c)
                pl = torch.zeros(4, n, h + 1, h + 1, o, device='cuda')
                for rr in (0, 1):
                    for cc in (0, 1):
                        gemm.igemm(A, B, n, h + 1, h + 1, o, gemm.up2_parity_taps(rr, cc), precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n,
                                   b_rows_per_tap=9 * o, out_f32=pl[rr * 2 + cc])
                t = torch.zeros(n, 2 * h + 2, 2 * h + 2, o, device='cuda', dtype=torch.float64)
                for rr in (0, 1):
                    for cc in (0, 1):
                        t[:, rr::2, cc::2] = pl[rr * 2 + cc].double()
                out = t[:, :2 * h + 1, :2 * h + 1]
                ref = F.conv_transpose2d(x.double(), wt.double().transpose(0, 1), stride=2).permute(0, 2, 3, 1)
            d = out.double() - ref
            print(f'{kind} c{c} {r}x{r}: rms {(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item():.2e}  max {(d.abs().max() / ref.abs().max()).item():.2e}', flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'small':
    small_res()


def x1_vs_x3():
    n = 16
    for (c, o, r) in [(32, 32, 1024), (64, 64, 512), (64, 32, 512), (128, 128, 256)]:
        for x3 in (True, False):
            x, wt, run = conv_case(n, c, o, r, r, x3, acc_k=512 if x3 else 0)
            cfg(0, 2)
            ms = bench_one(run)
            print(f'c{c} o{o} {r}x{r} x3={x3}: {ms:.3f} ms', flush=True)
            del x, wt, run
            torch.cuda.empty_cache()


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'x1x3':
    x1_vs_x3()


def gemm_probe():
    for (m, k, nn) in [(6400, 768, 2304), (6400, 768, 768), (6400, 768, 3072), (6400, 3072, 768), (3200, 3072, 768), (1600, 768, 2304)]:
        g = torch.Generator(device='cuda').manual_seed(1)
        a = torch.randn(m, k, device='cuda', generator=g)
        b = torch.randn(nn, k, device='cuda', generator=g) * 0.03
        A = gemm.split_planes(a, True).view(2, 1, m, k)
        B = gemm.split_planes(b * 64.0, True).view(2 * nn, k)
        out = torch.empty(m, nn, device='cuda')
        ref = a.double() @ b.double().t()
        row = {}
        for mode in (0, 1):
            cfg(0, mode)
            run = lambda: gemm.igemm(A, B, 1, 1, m, nn, gemm.TAPS_1X1, precision='x3', gain=1.0 / 64.0, acc_chunk_k=512, out_f32=out)
            ms = bench_one(run, iters=10)
            err = rel(out, ref)
            row[mode] = f'{ms * 1e3:.0f} us ({2.0 * m * k * nn / ms / 1e9:.0f} TF/s alg) err {err:.1e}'
        print(f'gemm {m}x{k}x{nn}: old {row[0]}   new {row[1]}', flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'gemm':
    gemm_probe()


def epilogue_cost():
    import math
    n = 16
    for (c, o, r) in [(128, 128, 256), (64, 64, 512), (32, 32, 1024)]:
        g = torch.Generator(device='cuda').manual_seed(5)
        x = torch.randn(n, c, r, r, device='cuda', generator=g)
        wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
        d = torch.rand(n, o, device='cuda', generator=g) + 0.5
        ps = torch.rand(n, o, device='cuda', generator=g) + 0.5
        bias = torch.randn(o, device='cuda', generator=g) * 0.1
        noise = torch.randn(r, r, device='cuda', generator=g) * 0.1
        rgb_w = torch.randn(n, 3, o, device='cuda', generator=g) * 0.1
        A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(-1, r, r, c)
        B = gemm.split_planes(wt.permute(2, 3, 0, 1).reshape(9 * o, c).contiguous(), True).reshape(-1, c)
        del x
        y = torch.zeros(2, n, r, r, o, device='cuda', dtype=torch.float16)
        xs = torch.zeros(2, n, r, r, o, device='cuda', dtype=torch.float16)
        o32 = torch.zeros(n, r, r, o, device='cuda')
        acc = torch.zeros(n, 3, r, r, device='cuda')
        base = dict(precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n, b_rows_per_tap=9 * o)
        act = dict(row_scale=d, bias=bias, noise=noise, noise_strides=(r, 1), act=1, alpha=0.2, gain=math.sqrt(2), clamp=256.0)
        variants = {
            'f32 only': dict(out_f32=o32),
            'act + y(hi,lo)': dict(out_hi=y[0], out_lo=y[1], **act),
            'act + xs(hi,lo) post': dict(out_hi=xs[0], out_lo=xs[1], post_scale=ps, **act),
            'act + xs + rgb': dict(out_hi=xs[0], out_lo=xs[1], post_scale=ps, rgb_w=rgb_w, rgb_acc=acc, **act),
            'act + y + xs + rgb (grad fwd)': dict(out_raw=y[0], out_raw_lo=y[1], out_hi=xs[0], out_lo=xs[1], post_scale=ps, rgb_w=rgb_w, rgb_acc=acc, **act),
            'act + rgb only': dict(rgb_w=rgb_w, rgb_acc=acc, **act),
        }
        cfg(0, 2)
        for name, kw in variants.items():
            ms = bench_one(lambda: gemm.igemm(A, B, n, r, r, o, gemm.TAPS_3X3, **base, **kw))
            print(f'c{c} {r}x{r} n{n} {name}: {ms:.3f} ms', flush=True)
        del A, B, y, xs, o32, acc
        torch.cuda.empty_cache()


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'epi':
    epilogue_cost()


def one_gradfwd(c, o, r, n=16):
    import math
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(n, c, r, r, device='cuda', generator=g)
    wt = torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05
    d = torch.rand(n, o, device='cuda', generator=g) + 0.5
    ps = torch.rand(n, o, device='cuda', generator=g) + 0.5
    bias = torch.randn(o, device='cuda', generator=g) * 0.1
    noise = torch.randn(r, r, device='cuda', generator=g) * 0.1
    rgb_w = torch.randn(n, 3, o, device='cuda', generator=g) * 0.1
    A = gemm.split_planes(x.permute(0, 2, 3, 1).contiguous(), True).reshape(-1, r, r, c)
    B = gemm.split_planes(wt.permute(2, 3, 0, 1).reshape(9 * o, c).contiguous(), True).reshape(-1, c)
    y = torch.zeros(2, n, r, r, o, device='cuda', dtype=torch.float16)
    xs = torch.zeros(2, n, r, r, o, device='cuda', dtype=torch.float16)
    acc = torch.zeros(n, 3, r, r, device='cuda')
    cfg(0, 2)
    for _ in range(3):
        gemm.igemm(A, B, n, r, r, o, gemm.TAPS_3X3, precision='x3', acc_chunk_k=512, a_plane_stride_imgs=n, b_rows_per_tap=9 * o,
                   row_scale=d, bias=bias, noise=noise, noise_strides=(r, 1), act=1, alpha=0.2, gain=math.sqrt(2), clamp=256.0,
                   out_raw=y[0], out_raw_lo=y[1], out_hi=xs[0], out_lo=xs[1], post_scale=ps, rgb_w=rgb_w, rgb_acc=acc)
    torch.cuda.synchronize()


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'onegrad':
    one_gradfwd(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
