"""GPU parity of the op-level API (stylemc_b200.ops) against the reference's own outputs (tests/golden/ops.npz, written by
oracle/pin_reference.py from torch_utils.ops impl='ref') and against the CPU oracle on other shapes."""
import numpy as np
import pytest
import torch

from oracle import act as o_act
from oracle import conv as o_conv
from oracle import fir as o_fir

pytestmark = pytest.mark.gpu

from cases import CONV_KW, FIR_KW  # noqa: E402


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to('cuda', dtype)


@pytest.mark.parametrize('name', list(FIR_KW))
def test_upfirdn2d_golden(golden, name):
    from stylemc_b200.ops import upfirdn2d
    g = golden('ops')
    f = dev(g[name + '.f']) if name + '.f' in g else None
    y = upfirdn2d.upfirdn2d(dev(g[name + '.x']), f, **FIR_KW[name])
    ref = torch.as_tensor(g[name + '.y'])
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max().item() <= 2e-6   # fp32 accumulate, different summation order than F.conv2d


def test_upsample2d_golden(golden):
    from stylemc_b200.ops import upfirdn2d
    g = golden('ops')
    f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
    y = upfirdn2d.upsample2d(dev(g['upsample2d.x']), f)
    assert (y.cpu() - torch.as_tensor(g['upsample2d.y'])).abs().max().item() <= 2e-6


@pytest.mark.parametrize('shape,kw', [
    ((3, 7, 129, 129), dict(padding=[1, 1, 1, 1], gain=4)),                  # conv0 tail at 64 px, ragged tile edges
    ((2, 3, 100, 60), dict(up=2, padding=[2, 1, 2, 1], gain=4)),             # skip-image upsample, non-square
    ((2, 3, 128, 128), dict(down=2, padding=[1, 2, 1, 2], flip_filter=True, gain=4)),
    ((1, 1, 1, 1), dict(up=2, padding=[2, 1, 2, 1], gain=4)),                # smallest legal input
])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.float64])
def test_upfirdn2d_vs_oracle(shape, kw, dtype):
    from stylemc_b200.ops import upfirdn2d
    g = torch.Generator().manual_seed(0)
    x = torch.randn(*shape, generator=g)
    f = o_fir.setup_filter([1, 3, 3, 1])
    ref = o_fir.upfirdn2d(x.double(), f.double(), **kw)
    y = upfirdn2d.upfirdn2d(x.to('cuda', dtype), f.cuda(), **kw)
    tol = {torch.float32: 2e-6, torch.float16: 4e-3, torch.float64: 1e-7}[dtype]    # f is float32 in the plugin ABI
    assert y.dtype == dtype
    assert (y.double().cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


def test_upfirdn2d_channels_last_and_grad():
    from stylemc_b200.ops import upfirdn2d
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 8, 20, 20, generator=g)
    f = o_fir.setup_filter([1, 3, 3, 1])
    kw = dict(up=2, padding=[2, 1, 2, 1], gain=4)
    xr = x.clone().requires_grad_(True)
    yr = o_fir.upfirdn2d(xr, f, **kw)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = upfirdn2d.upfirdn2d(xc, f.cuda(), **kw)
    y.backward(dy.cuda())
    assert (y.detach().cpu() - yr.detach()).abs().max().item() <= 2e-6
    assert (xc.grad.cpu() - xr.grad).abs().max().item() <= 5e-6


def test_upfirdn2d_errors():
    from stylemc_b200.ops import upfirdn2d
    f = o_fir.setup_filter([1, 3, 3, 1]).cuda()
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(torch.zeros(1, 1, 4, 4), f)                      # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(torch.zeros(1, 1, 2, 2, device='cuda'), f, padding=-2)   # output smaller than 1x1
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(torch.zeros(1, 1, 4, 4, device='cuda'), f, impl='ref')


@pytest.mark.parametrize('name', list(o_act.ACTIVATIONS))
@pytest.mark.parametrize('tag,kw', [('def', {}), ('clamp', dict(gain=1.7, clamp=0.9, alpha=0.3))])
def test_bias_act_golden(golden, name, tag, kw):
    from stylemc_b200.ops import bias_act
    g = golden('ops')
    y = bias_act.bias_act(dev(g['bias_act.x']), dev(g['bias_act.b']), act=name, **kw)
    ref = torch.as_tensor(g[f'bias_act.{name}.{tag}'])
    assert (y.cpu() - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('kw', [dict(act='lrelu', gain=2 ** 0.5, clamp=256 * 2 ** 0.5), dict(act='linear', clamp=256),
                                dict(act='lrelu', gain=1.5, clamp=1.0), dict(act='swish'), dict(act='tanh', clamp=0.5)])
@pytest.mark.parametrize('shape', [(3, 6, 5, 4), (2, 32, 64, 64), (1, 5, 7, 3)])
def test_bias_act_fwd_bwd_vs_oracle(kw, shape):
    from stylemc_b200.ops import bias_act
    g = torch.Generator().manual_seed(2)
    x = torch.randn(*shape, generator=g) * 60
    b = torch.randn(shape[1], generator=g)
    dy = torch.randn(*shape, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = o_act.bias_act(xr, b, **kw)
    yr.backward(dy)
    xc = x.cuda().requires_grad_(True)
    y = bias_act.bias_act(xc, b.cuda(), **kw)
    y.backward(dy.cuda())
    assert (y.detach().cpu() - yr.detach()).abs().max().item() <= 2e-6 * max(1.0, yr.abs().max().item())
    assert (xc.grad.cpu() - xr.grad).abs().max().item() <= 2e-6 * max(1.0, xr.grad.abs().max().item())


@pytest.mark.parametrize('shape', [(3, 10, 33, 7), (2, 6, 8, 24), (5, 64, 40, 40)])
def test_bias_act_fp16_vectors(shape):
    """fp16 fast path: two 8-element vectors per thread and iteration, per-vector or per-element bias lookup (H*W % 8), ragged tail."""
    from stylemc_b200.ops import bias_act
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(*shape, generator=g) * 3).half()
    b = torch.randn(shape[1], generator=g).half()
    dy = torch.randn(*shape, generator=g).half()
    kw = dict(act='lrelu', gain=2 ** 0.5, clamp=4.0)
    xr = x.float().requires_grad_(True)
    yr = o_act.bias_act(xr, b.float(), **kw)
    yr.backward(dy.float())
    xc = x.cuda().requires_grad_(True)
    y = bias_act.bias_act(xc, b.cuda(), **kw)
    y.backward(dy.cuda())
    assert y.dtype == torch.float16
    assert (y.detach().float().cpu() - yr.detach()).abs().max().item() <= 4e-3      # one fp16 rounding of |y| <= 4
    # the clamp mask is taken on the STORED (fp16-rounded) y, as in the reference plugin: skip outputs within a rounding step of the clamp
    ok = ~((yr.detach().abs() < 4.0) & (yr.detach().abs() > 4.0 - 4e-3))
    assert ((xc.grad.float().cpu() - xr.grad).abs() * ok).max().item() <= 4e-3 * max(1.0, xr.grad.abs().max().item())
    assert ok.float().mean().item() > 0.99


def test_bias_act_misc():
    from stylemc_b200.ops import bias_act
    g = torch.Generator().manual_seed(3)
    x2, b2 = torch.randn(4, 7, generator=g), torch.randn(4, generator=g)
    y = bias_act.bias_act(x2.cuda(), b2.cuda(), dim=0, act='lrelu')
    assert (y.cpu() - o_act.bias_act(x2, b2, dim=0, act='lrelu')).abs().max().item() <= 1e-6
    xh = torch.randn(2, 16, 9, 9, generator=g)
    yh = bias_act.bias_act(xh.cuda().half().contiguous(memory_format=torch.channels_last), b=None, act='lrelu', clamp=1.0)
    assert yh.dtype == torch.float16 and yh.stride(1) == 1
    assert (yh.float().cpu() - o_act.bias_act(xh.half().float(), act='lrelu', clamp=1.0)).abs().max().item() <= 2e-3
    assert bias_act.bias_act(torch.zeros(0, 3, 2, 2, device='cuda'), torch.zeros(3, device='cuda')).numel() == 0   # empty input
    with pytest.raises(RuntimeError):
        bias_act.bias_act(torch.zeros(2, 3), torch.zeros(3))                  # CPU tensors
    with pytest.raises(RuntimeError):
        bias_act.bias_act(torch.zeros(2, 3, device='cuda'), torch.zeros(4, device='cuda'))   # wrong bias size


@pytest.mark.parametrize('name', list(CONV_KW))
def test_conv2d_resample_golden(golden, name):
    from stylemc_b200.ops import conv2d_resample, upfirdn2d
    g = golden('ops')
    kw = CONV_KW[name]
    f = upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda') if kw.get('up', 1) > 1 else None
    y = conv2d_resample.conv2d_resample(dev(g[name + '.x']), dev(g[name + '.w']), f=f, **kw)
    ref = torch.as_tensor(g[name + '.y'])
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())   # split-fp16 operands ~2^-21


def test_conv2d_resample_input_grad():
    from stylemc_b200.ops import conv2d_gradfix, conv2d_resample, upfirdn2d
    g = torch.Generator().manual_seed(4)
    f = o_fir.setup_filter([1, 3, 3, 1])
    for kw, xs, ws in ((dict(padding=1), (2, 40, 9, 9), (33, 40, 3, 3)),
                       (dict(up=2, padding=1, flip_weight=False), (2, 16, 6, 6), (8, 16, 3, 3)),
                       (dict(), (2, 40, 5, 5), (3, 40, 1, 1))):
        x, w = torch.randn(*xs, generator=g), torch.randn(*ws, generator=g)
        ff = f if kw.get('up', 1) > 1 else None
        xr = x.clone().requires_grad_(True)
        yr = o_conv.conv2d_resample(xr, w, f=ff, **kw)
        dy = torch.randn(yr.shape, generator=g)
        yr.backward(dy)
        xc = x.cuda().requires_grad_(True)
        with conv2d_gradfix.no_weight_gradients():
            y = conv2d_resample.conv2d_resample(xc, w.cuda(), f=None if ff is None else ff.cuda(), **kw)
            y.backward(dy.cuda())
        assert (y.detach().cpu() - yr.detach()).abs().max().item() <= 2e-5 * yr.abs().max().item()
        assert (xc.grad.cpu() - xr.grad).abs().max().item() <= 2e-5 * xr.grad.abs().max().item()


def test_fma_golden(golden):
    """ops.fma (smc_fma / smc_fma_reduce) against the reference's own autograd.Function run on the CPU (fma.py:15-58): forward and the
    three broadcast-aware input gradients."""
    from stylemc_b200.ops import fma
    g = golden('ops')
    a, b, c = (dev(g['fma.' + k]).requires_grad_(True) for k in 'abc')
    y = fma.fma(a, b, c)
    assert (y.detach().cpu() - torch.as_tensor(g['fma.y'])).abs().max().item() <= 1e-6
    y.backward(dev(g['fma.dy']))
    for t, k in ((a, 'da'), (b, 'db'), (c, 'dc')):
        ref = torch.as_tensor(g['fma.' + k])
        assert t.grad.shape == ref.shape
        assert (t.grad.cpu() - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item()), k


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.float64])
def test_fma_vs_oracle_shapes_and_second_order(dtype):
    """Long reductions (block-per-output kernel), scalars and rank < 4 operands; fp16 / fp64; gradient of the gradient."""
    gen = torch.Generator().manual_seed(11)
    from stylemc_b200.ops import fma
    tol = {torch.float32: 2e-6, torch.float16: 2e-3, torch.float64: 1e-12}[dtype]
    for sa, sb, sc in (((3, 5, 40, 40), (3, 5, 1, 1), (40, 40)), ((2, 1, 7), (4, 1), (1,)), ((6,), (), (6,)), ((2, 3, 4, 4), (2, 3, 4, 4), (1, 3, 1, 1))):
        ta, tb, tc = (torch.randn(s, generator=gen, dtype=torch.float64) for s in (sa, sb, sc))
        ra, rb, rc = (t.clone().requires_grad_(True) for t in (ta, tb, tc))
        yr = o_conv.fma(ra, rb, rc)
        dy = torch.randn(yr.shape, generator=gen, dtype=torch.float64)
        gr = torch.autograd.grad(yr, (ra, rb, rc), dy)
        xa, xb, xc = (t.to(dtype).cuda().requires_grad_(True) for t in (ta, tb, tc))
        y = fma.fma(xa, xb, xc)
        assert y.dtype == dtype and y.shape == yr.shape
        assert (y.detach().cpu().double() - yr.detach()).abs().max().item() <= tol * max(1.0, yr.abs().max().item())
        gx = torch.autograd.grad(y, (xa, xb, xc), dy.to(dtype).cuda(), create_graph=True)
        for got, want in zip(gx, gr):
            assert got.shape == want.shape
            assert (got.detach().cpu().double() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item()) * 40
        if dtype == torch.float64:                      # d/da of sum(d y/d b * v) = v broadcast * dy: second order through _FmaReduce.backward
            v = torch.randn(tb.shape, generator=gen, dtype=torch.float64)
            ga2, = torch.autograd.grad((gx[1] * v.cuda()).sum(), xa)
            ra2 = ta.clone().requires_grad_(True)
            gb_ref, = torch.autograd.grad(o_conv.fma(ra2, rb, rc), rb, dy, create_graph=True)
            want, = torch.autograd.grad((gb_ref * v).sum(), ra2)
            assert (ga2.cpu() - want).abs().max().item() <= 1e-10


def test_conv2d_gradfix_entry_points_golden(golden):
    """conv2d_gradfix.conv2d / conv_transpose2d (conv2d_gradfix.py:35-43) against the reference's own functions run on the CPU
    (F.conv2d / F.conv_transpose2d with bias), forward and the input gradient; weight gradients are refused outside
    no_weight_gradients() and the switch is restored when the block raises."""
    from stylemc_b200.ops import conv2d_gradfix
    g = golden('ops')
    x, w, b, wt = dev(g['gradfix.x']), dev(g['gradfix.w']), dev(g['gradfix.bias']), dev(g['gradfix.wt'])
    for got, key in ((conv2d_gradfix.conv2d(x, w, bias=b, padding=1), 'gradfix.conv2d'),
                     (conv2d_gradfix.conv_transpose2d(x, wt, stride=2), 'gradfix.conv_transpose2d')):
        ref = torch.as_tensor(g[key])
        assert got.shape == ref.shape
        assert (got.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), key
    xg = x.clone().requires_grad_(True)
    with conv2d_gradfix.no_weight_gradients():
        conv2d_gradfix.conv2d(xg, w, padding=1).backward(dev(g['gradfix.dy']))
    ref = torch.as_tensor(g['gradfix.dx'])
    assert (xg.grad.cpu() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    wg = w.clone().requires_grad_(True)
    with pytest.raises(RuntimeError):
        conv2d_gradfix.conv2d(x, wg, padding=1).sum().backward()          # weight gradient outside no_weight_gradients()
    with pytest.raises(ZeroDivisionError):
        with conv2d_gradfix.no_weight_gradients():
            1 / 0
    assert conv2d_gradfix.weight_gradients_disabled is False
    with pytest.raises(RuntimeError):
        conv2d_gradfix.conv2d(x, w, stride=2)


def test_mask_scale_kernel():
    """smc_mask_scale: g * mask * scale (the ToRGB clamp mask and the loss scale on the incoming image gradient) == the two ATen multiplies."""
    from stylemc_b200 import _lib
    gen = torch.Generator().manual_seed(11)
    for shape in ((2, 3, 64, 64), (1, 3, 5, 7)):                       # vector path / scalar tail path
        g = torch.randn(shape, generator=gen).cuda()
        m = (torch.rand(shape, generator=gen) > 0.3).to(torch.uint8).cuda()
        k = torch.tensor([64.0], device='cuda')
        out = torch.empty_like(g)
        _lib.call('smc_mask_scale', _lib.ptr(g), _lib.ptr(m), _lib.ptr(k), _lib.ptr(out), g.numel(), _lib.stream())
        assert torch.equal(out, (g * m) * k)
