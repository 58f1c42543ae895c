"""GPU checks of the tcgen05 implicit-GEMM kernel (smc_igemm) against plain torch fp32/fp64 references of the same
contraction (a floating-point kernel: tolerance stated per case)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def nhwc_half(x):
    return x.permute(0, 2, 3, 1).contiguous().half()


def tap_matrix(w):
    """[O, I, kh, kw] -> [kh*kw*O, I] (row t*O + o)."""
    o, i, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw * o, i).contiguous()


@pytest.mark.parametrize('m,k,n', [(300, 768, 2304), (128, 64, 32), (50, 3072, 768), (1000, 96, 96)])
def test_plain_gemm_x1(m, k, n):
    from stylemc_b200 import gemm
    g = torch.Generator(device='cuda').manual_seed(0)
    a = torch.randn(m, k, device='cuda', generator=g).half()
    b = torch.randn(n, k, device='cuda', generator=g).half()
    out = torch.empty(m, n, device='cuda')
    gemm.igemm(a.view(1, 1, m, k), b, 1, 1, m, n, gemm.TAPS_1X1, out_f32=out)
    ref = a.double() @ b.double().t()
    assert rel(out, ref) <= 1e-5          # fp16 products are exact in fp32; only the accumulation order differs


def test_plain_gemm_x3_is_fp32_grade():
    from stylemc_b200 import gemm
    g = torch.Generator(device='cuda').manual_seed(1)
    m, k, n = 200, 768, 768
    a = torch.randn(m, k, device='cuda', generator=g)
    b = torch.randn(n, k, device='cuda', generator=g) * 0.03
    A = gemm.split_planes(a, True).view(2, 1, m, k)
    # small weights: scale by a power of two so the lo plane stays out of the fp16 subnormal range, undo it in the epilogue
    B = gemm.split_planes(b * 64.0, True).view(2 * n, k)
    out = torch.empty(m, n, device='cuda')
    gemm.igemm(A, B, 1, 1, m, n, gemm.TAPS_1X1, precision='x3', gain=1.0 / 64.0, out_f32=out)
    ref = a.double() @ b.double().t()
    # limited by the tensor core's fp32 accumulation (truncating adds over K = 768), not by the operand split
    assert rel(out, ref) <= 1e-5


@pytest.mark.parametrize('n,c,o,h,w', [(2, 64, 64, 16, 16), (3, 512, 512, 4, 4), (1, 32, 32, 40, 24), (2, 128, 256, 33, 17), (5, 64, 32, 8, 8)])
def test_conv3x3(n, c, o, h, w):
    from stylemc_b200 import gemm
    g = torch.Generator(device='cuda').manual_seed(2)
    x = torch.randn(n, c, h, w, device='cuda', generator=g).half().float()
    wt = (torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.05).half().float()
    out = torch.empty(n, h, w, o, device='cuda')
    gemm.igemm(nhwc_half(x), tap_matrix(wt).half(), n, h, w, o, gemm.TAPS_3X3, out_f32=out)
    ref = F.conv2d(x.double(), wt.double(), padding=1).permute(0, 2, 3, 1)
    assert rel(out, ref) <= 1e-5


def test_epilogue_all_outputs():
    from stylemc_b200 import gemm
    g = torch.Generator(device='cuda').manual_seed(3)
    n, c, o, h, w = 3, 64, 128, 16, 16
    x = torch.randn(n, c, h, w, device='cuda', generator=g).half().float()
    wt = (torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.1).half().float()
    d = torch.rand(n, o, device='cuda', generator=g) + 0.5
    s = torch.randn(n, o, device='cuda', generator=g)
    bias = torch.randn(o, device='cuda', generator=g)
    noise = torch.randn(h, w, device='cuda', generator=g)
    res = torch.randn(n, h, w, o, device='cuda', generator=g)
    o32 = torch.empty(n, h, w, o, device='cuda')
    ohi, olo, oraw = (torch.empty(n, h, w, o, device='cuda', dtype=torch.float16) for _ in range(3))
    gemm.igemm(nhwc_half(x), tap_matrix(wt).half(), n, h, w, o, gemm.TAPS_3X3, row_scale=d, post_scale=s, bias=bias, noise=noise,
               noise_strides=(w, 1), act=1, alpha=0.2, gain=2 ** 0.5, clamp=3.0, residual=res, out_f32=o32, out_hi=ohi, out_lo=olo,
               out_raw=oraw)
    u = F.conv2d(x.double(), wt.double(), padding=1)
    z = u * d.double()[:, :, None, None] + noise.double() + bias.double()[None, :, None, None]
    y = (F.leaky_relu(z, 0.2) * 2 ** 0.5).clamp(-3, 3).permute(0, 2, 3, 1)
    full = y * s.double()[:, None, None, :] + res.double()
    assert rel(oraw.float(), y) <= 1e-3
    assert rel(o32, full) <= 1e-5
    assert rel(ohi.float() + olo.float(), full) <= 1e-5
    assert rel(ohi.float(), full) <= 1e-3


def test_transposed_stride2_parity_planes():
    from stylemc_b200 import gemm
    g = torch.Generator(device='cuda').manual_seed(4)
    n, c, o, h, w = 2, 64, 32, 8, 8
    x = torch.randn(n, c, h, w, device='cuda', generator=g).half().float()
    wt = (torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.1).half().float()
    planes = torch.zeros(4, n, h + 1, w + 1, o, device='cuda')
    for r in (0, 1):
        for cc in (0, 1):
            gemm.igemm(nhwc_half(x), tap_matrix(wt).half(), n, h + 1, w + 1, o, gemm.up2_parity_taps(r, cc), out_f32=planes[r * 2 + cc])
    ref = F.conv_transpose2d(x.double(), wt.double().transpose(0, 1), stride=2)            # [n, o, 2h+1, 2w+1]
    t = torch.zeros(n, 2 * h + 2, 2 * w + 2, o, device='cuda', dtype=torch.float64)
    for r in (0, 1):
        for cc in (0, 1):
            t[:, r::2, cc::2] = planes[r * 2 + cc].double()
    assert rel(t[:, :2 * h + 1, :2 * w + 1], ref.permute(0, 2, 3, 1)) <= 1e-5
    assert t[:, 2 * h + 1].abs().max().item() == 0 and t[:, :, 2 * w + 1].abs().max().item() == 0   # cells outside the grid are zero

    # dgrad of the same conv from gradient parity planes
    gy = torch.randn(n, o, 2 * h + 1, 2 * w + 1, device='cuda', generator=g).half().float()
    gp = torch.zeros(4, n, h + 1, w + 1, o, device='cuda', dtype=torch.float16)
    gyl = F.pad(gy, (0, 1, 0, 1)).permute(0, 2, 3, 1)
    for r in (0, 1):
        for cc in (0, 1):
            gp[r * 2 + cc] = gyl[:, r::2, cc::2].half()
    bt = wt.permute(2, 3, 1, 0).reshape(9 * c, o).contiguous().half()                      # rows t*Cin + i, cols o
    gx = torch.empty(n, h, w, c, device='cuda')
    gemm.igemm(gp.view(4 * n, h + 1, w + 1, o), bt, n, h, w, c, gemm.up2_dgrad_taps(n), out_f32=gx)
    xr = x.double().requires_grad_(True)
    F.conv_transpose2d(xr, wt.double().transpose(0, 1), stride=2).backward(gy.double())
    assert rel(gx, xr.grad.permute(0, 2, 3, 1)) <= 1e-5


def test_conv3x3_dgrad_taps():
    from stylemc_b200 import gemm
    g = torch.Generator(device='cuda').manual_seed(5)
    n, c, o, h, w = 2, 64, 96, 12, 12
    wt = (torch.randn(o, c, 3, 3, device='cuda', generator=g) * 0.1).half().float()
    gy = torch.randn(n, o, h, w, device='cuda', generator=g).half().float()
    bt = wt.permute(2, 3, 1, 0).reshape(9 * c, o).contiguous().half()
    gx = torch.empty(n, h, w, c, device='cuda')
    gemm.igemm(nhwc_half(gy), bt, n, h, w, c, gemm.TAPS_3X3_DGRAD, out_f32=gx)
    xr = torch.zeros(n, c, h, w, device='cuda', dtype=torch.float64, requires_grad=True)
    F.conv2d(xr, wt.double(), padding=1).backward(gy.double())
    assert rel(gx, xr.grad.permute(0, 2, 3, 1)) <= 1e-5


def test_argument_errors():
    from stylemc_b200 import gemm
    a = torch.zeros(1, 1, 128, 48, device='cuda', dtype=torch.float16)     # K not a multiple of 32
    b = torch.zeros(32, 48, device='cuda', dtype=torch.float16)
    with pytest.raises(RuntimeError):
        gemm.igemm(a, b, 1, 1, 128, 32, gemm.TAPS_1X1, out_f32=torch.empty(128, 32, device='cuda'))


@pytest.mark.parametrize('shape,pad', [((64, 32, 3, 3), 1), ((33, 40, 3, 3), 32), ((3, 6, 1, 1), 32), ((768, 3072), 1), ((512, 512, 3, 3), 1)])
def test_prepare_weights_matches_the_torch_formulation_bitwise(shape, pad):
    """smc_prepare_weights (one launch) against the ATen chain it replaces: permute -> reshape -> half / sub / half (gemm.split_planes),
    square().sum(); planes bit-identical, q to fp32 rounding; with the device-side power-of-two prescale of the CLIP linears."""
    from stylemc_b200 import gemm
    w = (torch.randn(*shape, generator=torch.Generator().manual_seed(5)) * 0.03).cuda()
    o, i = shape[:2]
    t = shape[2] * shape[3] if len(shape) == 4 else 1
    op, ip = -(-o // pad) * pad, -(-i // pad) * pad
    for prescale in (False, True):
        Bf, Bb, q, inv = gemm.prepare_weights(w, two=True, fwd=True, bwd=True, q=True, prescale=prescale, pad_to=pad)
        k = gemm.pow2_prescale(w) if prescale else 1.0
        assert inv == 1.0 / k
        m = torch.zeros([t, op, ip], device='cuda')
        m[:, :o, :i] = (w * k).reshape(o, i, t).permute(2, 0, 1)
        assert torch.equal(Bf, gemm.split_planes(m.reshape(t * op, ip), True).reshape(-1, ip))
        assert torch.equal(Bb, gemm.split_planes(m.permute(0, 2, 1).reshape(t * ip, op).contiguous(), True).reshape(-1, op))
        qr = w.reshape(o, i, t).double().square().sum(2)
        assert ((q.double() - qr).abs() <= 1e-6 * qr.abs().max()).all()
    B1, none_b, none_q, _ = gemm.prepare_weights(w, two=False)
    assert none_b is None and none_q is None and torch.equal(B1, gemm.prepare_weights(w, two=True)[0][:t * o])     # hi plane alone
