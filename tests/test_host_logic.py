"""CPU: host-side logic of the product (no kernel launches): resampling tap tables, implicit-GEMM tap tables (checked by
emulating the smc_igemm contract with plain torch), S-space bookkeeping, network parameter naming, schedules."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import direction as o_dir
from oracle import synthesis as o_syn


def emulate_igemm(A, B, n_img, H, W, n_out, taps):
    """D[n,h,w,o] = sum_t sum_c A[n+dn_t, h+dy_t, w+dx_t, c] * B[tap_t*n_out + o, c], zero outside A (include/stylemc_b200.h)."""
    NA, HA, WA, C = A.shape
    out = torch.zeros(n_img, H, W, n_out, dtype=A.dtype)
    for t in taps:
        dn, dy, dx, ti = t if len(t) == 4 else (0,) + tuple(t)
        shifted = torch.zeros(n_img, H, W, C, dtype=A.dtype)
        for n in range(n_img):
            if not 0 <= n + dn < NA:
                continue
            h0, h1 = max(0, -dy), min(H, HA - dy)
            w0, w1 = max(0, -dx), min(W, WA - dx)
            if h1 > h0 and w1 > w0:
                shifted[n, h0:h1, w0:w1] = A[n + dn, h0 + dy:h1 + dy, w0 + dx:w1 + dx]
        out += shifted @ B[ti * n_out:(ti + 1) * n_out].t()
    return out


def test_tap_tables_reproduce_the_convolutions():
    from stylemc_b200 import gemm
    g = torch.Generator().manual_seed(0)
    n, c, o, h = 2, 5, 4, 6
    x = torch.randn(n, c, h, h, generator=g, dtype=torch.float64)
    w = torch.randn(o, c, 3, 3, generator=g, dtype=torch.float64)
    A = x.permute(0, 2, 3, 1).contiguous()
    Bf = w.permute(2, 3, 0, 1).reshape(9 * o, c)
    Bb = w.permute(2, 3, 1, 0).reshape(9 * c, o)
    # forward 3x3
    y = emulate_igemm(A, Bf, n, h, h, o, gemm.TAPS_3X3)
    assert torch.allclose(y, F.conv2d(x, w, padding=1).permute(0, 2, 3, 1), atol=1e-12)
    # its dgrad
    gy = torch.randn(n, o, h, h, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, w, padding=1).backward(gy)
    gx = emulate_igemm(gy.permute(0, 2, 3, 1).contiguous(), Bb, n, h, h, c, gemm.TAPS_3X3_DGRAD)
    assert torch.allclose(gx, xr.grad.permute(0, 2, 3, 1), atol=1e-12)
    # transposed stride-2 conv as four parity GEMMs
    ref = F.conv_transpose2d(x, w.transpose(0, 1), stride=2).permute(0, 2, 3, 1)            # [n, 2h+1, 2h+1, o]
    t = torch.zeros(n, 2 * h + 2, 2 * h + 2, o, dtype=torch.float64)
    for r in (0, 1):
        for cc in (0, 1):
            t[:, r::2, cc::2] = emulate_igemm(A, Bf, n, h + 1, h + 1, o, gemm.up2_parity_taps(r, cc))
    assert torch.allclose(t[:, :2 * h + 1, :2 * h + 1], ref, atol=1e-12)
    assert t[:, 2 * h + 1].abs().max() == 0 and t[:, :, 2 * h + 1].abs().max() == 0
    # and its dgrad from gradient parity planes stacked on the image axis
    gt = torch.randn(n, o, 2 * h + 1, 2 * h + 1, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    F.conv_transpose2d(xr, w.transpose(0, 1), stride=2).backward(gt)
    gl = F.pad(gt, (0, 1, 0, 1)).permute(0, 2, 3, 1)
    gp = torch.cat([gl[:, r::2, cc::2] for r in (0, 1) for cc in (0, 1)])                    # [4n, h+1, h+1, o]
    gx = emulate_igemm(gp, Bb, n, h, h, c, gemm.up2_dgrad_taps(n))
    assert torch.allclose(gx, xr.grad.permute(0, 2, 3, 1), atol=1e-12)


def test_x3_tap_expansion_counts():
    from stylemc_b200 import _lib, gemm
    assert len(gemm.TAPS_3X3) * 3 <= _lib.MAX_TAPS
    assert sorted(t[2] for r in (0, 1) for c in (0, 1) for t in gemm.up2_parity_taps(r, c)) == list(range(9))


@pytest.mark.parametrize('res', [64, 100, 224, 256, 512, 1024])
def test_antialias_tables_match_interpolate(res):
    from stylemc_b200 import resample
    x = torch.randn(1, 1, res, res, dtype=torch.float64, generator=torch.Generator().manual_seed(res))
    ref = F.interpolate(x, size=(224, 224), mode='bicubic', antialias=True, align_corners=False)[0, 0]
    M = torch.as_tensor(resample.dense_matrix(res, 224))
    assert (M @ x[0, 0] @ M.t() - ref).abs().max().item() <= 2e-6
    start, count, wgt, taps = resample.aa_tables(res, 224)
    oidx, cnt, wt, taps_t = resample.transpose_tables(start, count, wgt, res)
    Mt = torch.zeros(224, res, dtype=torch.float64)
    for i in range(res):
        for k in range(cnt[i]):
            Mt[oidx[i, k], i] += float(wt[i, k])
    assert torch.allclose(Mt, M.to(torch.float64), atol=1e-7)


def test_unprocess_constants_match_reference():
    from stylemc_b200 import resample
    assert resample.CLIP_MEAN == o_dir.CLIP_MEAN and resample.CLIP_STD == o_dir.CLIP_STD


def test_s_space_bookkeeping_matches_oracle():
    """split_ws / get_styles / get_temp_shapes (utils.py:77-158) on the product's own network modules."""
    from stylemc_b200 import networks, utils
    G = networks.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    Go = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == {k: tuple(v.shape) for k, v in Go.state_dict().items()}
    Go.load_state_dict(G.state_dict())
    ws = torch.randn(3, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(2))
    for a, b in zip(utils.split_ws(G, ws), o_syn.split_ws(Go, ws)):
        assert torch.equal(a, b)
    S, shapes = utils.get_styles(G, ws, utils.split_ws(G, ws), 'cpu')
    So, shapes_o = o_syn.get_styles(Go, ws, o_syn.split_ws(Go, ws))
    assert shapes == shapes_o and (S - So).abs().max().item() <= 1e-5
    assert S.shape == (3, 26, 512) and S[:, 15:].abs().max().item() == 0           # rows beyond this net stay zero
    assert utils.get_temp_shapes(networks.make_generator(64, seed=1, channel_base=2048, channel_max=512)) == shapes


def test_parameter_names_follow_legacy_map():
    """legacy.py:173-202 names: synthesis.b{res}.{conv0,conv1,torgb}.{weight,bias,affine.weight,affine.bias,noise_const,...}."""
    from stylemc_b200 import networks
    G = networks.make_generator(32, seed=0, channel_base=1024)
    sd = G.state_dict()
    for key in ('synthesis.b4.const', 'synthesis.b4.conv1.weight', 'synthesis.b4.conv1.affine.weight', 'synthesis.b4.conv1.noise_const',
                'synthesis.b4.conv1.noise_strength', 'synthesis.b4.torgb.weight', 'synthesis.b8.conv0.weight', 'synthesis.b8.conv0.resample_filter',
                'synthesis.b32.torgb.affine.bias', 'synthesis.b32.resample_filter'):
        assert key in sd, key
    assert tuple(sd['synthesis.b8.conv0.weight'].shape) == (128, 256, 3, 3) and tuple(sd['synthesis.b8.torgb.weight'].shape) == (3, 128, 1, 1)
    assert G.synthesis.num_ws == 2 * int(math.log2(32)) - 2
    assert all(not p.requires_grad for p in G.parameters())


def test_schedules_and_sharding():
    from stylemc_b200 import direction
    for it in (1, 5, 10):
        assert abs(direction.cosine_lr(1.5, it, 10) - o_dir.cosine_lr(1.5, it, 10)) < 1e-15
    assert direction.S_TRAINABLE_SPACE_CHANNELS == o_dir.S_TRAINABLE_ROWS and direction.RESOLUTION_DICT == o_dir.RESOLUTION_TO_K
    for n, world in ((129, 8), (64, 8), (5, 2), (3, 4)):
        spans = [direction.shard_rows(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_fma_has_no_cpu_path_and_plans_broadcast_strides():
    """ops.fma runs on repo kernels only (smc_fma / smc_fma_reduce): CPU tensors raise; the host-side broadcast bookkeeping (the role
    of fma.py:49-58 `_unbroadcast`) is checked here, the numbers on the GPU (tests/test_ops_gpu.py::test_fma_golden)."""
    from stylemc_b200.ops import fma
    a, b, c = torch.zeros(2, 3, 4, 4), torch.zeros(2, 3, 1, 1), torch.zeros(4, 4)
    with pytest.raises(RuntimeError):
        fma.fma(a, b, c)
    space = fma._space(a, b, c)
    assert space == (2, 3, 4, 4)
    assert fma._strides(a, space) == [48, 16, 4, 1] and fma._strides(b, space) == [3, 1, 0, 0] and fma._strides(c, space) == [0, 0, 4, 1]
    assert fma._space(torch.zeros(5), torch.zeros(3, 1)) == (1, 1, 3, 5)
    with pytest.raises(RuntimeError):
        fma._space(torch.zeros(1, 1, 1, 1, 2))


def test_setup_filter_matches_oracle():
    from oracle import fir
    from stylemc_b200.ops import upfirdn2d
    for taps, kw in (([1, 3, 3, 1], {}), ([1, 2, 3, 4, 4, 3, 2, 1], {}), ([1, 2, 1], dict(gain=3, flip_filter=True)), (None, {})):
        assert torch.equal(upfirdn2d.setup_filter(taps, **kw), fir.setup_filter(taps, **kw))


def test_bench_flop_accounting_matches_survey_8d():
    """bench.py's algorithmic work per image (the numerator of roofline.achieved) against SURVEY.md section 8(d): synthesis forward
    148.13 GFLOP at 1024 px, 90.14 GFLOP for the 256-px truncation, ViT-B/32 forward 8.82 GFLOP."""
    import importlib.util
    import os
    from types import SimpleNamespace as NS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_module', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ch = lambda r: min(32768 // r, 512)
    blocks = []
    for k, r in enumerate([4, 8, 16, 32, 64, 128, 256, 512, 1024]):
        conv0 = None if k == 0 else NS(cin=ch(r // 2), cout=ch(r))
        blocks.append(NS(resolution=r, conv0=conv0, conv1=NS(cin=ch(r), cout=ch(r)), torgb=NS(cin=ch(r))))
    assert abs(bench.synth_flops_per_image(blocks, 8) / 1e9 - 148.13) < 0.4          # + ToRGB 0.39 GFLOP
    assert abs(bench.synth_flops_per_image(blocks, 6) / 1e9 - 90.14) < 0.2
    assert abs(bench.VIT_FLOPS_FWD / 1e9 - 8.82) < 0.1
    p = bench.peaks()
    assert p['hbm'] > 1000 and p['tflops'] > 100


@pytest.mark.parametrize('towers,overlap', [(1, False), (2, False), (1, True), (2, True)])
def test_direction_finder_combines_towers_and_loss_scales(towers, overlap, monkeypatch):
    """Host logic of DirectionFinder.loss_and_grad with stand-in engines on the CPU (linear synthesis, linear towers): for
    clip_type='double' the loss is loss1 + 0.5 * loss2 (find_direction.py:164) and the pixel gradients of the two towers, each
    carrying its own power-of-two loss scale, are reconciled before the synthesis backward.  overlap=True walks the two-stream
    branch (the benchmark's default) with the CUDA stream calls replaced by no-ops."""
    import contextlib
    from stylemc_b200 import direction

    class FakeStream:
        def wait_stream(self, other):
            pass
    if overlap:
        monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: FakeStream())
        monkeypatch.setattr(torch.cuda, 'Stream', lambda device=None: FakeStream())
        monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
        monkeypatch.setattr(torch.Tensor, 'record_stream', lambda self, s: None)
    gen = torch.Generator().manual_seed(9)
    n, rows, coef = 3, direction.S_TRAINABLE_SPACE_CHANNELS, 0.7
    A = torch.randn(26 * 512, 3 * 8 * 8, generator=gen, dtype=torch.float64) * 0.02
    text = torch.nn.functional.normalize(torch.randn(1, 16, generator=gen, dtype=torch.float64), dim=1)

    class Engine:
        def forward(self, s, until_k, noise_mode, save, grad_rows=None):
            return None, (s.double().reshape(len(s), -1) @ A).reshape(len(s), 3, 8, 8), s

        def backward(self, saved, g_img, rows_, noise_mode):
            return (g_img.reshape(len(g_img), -1) @ A.t()).reshape(len(g_img), 26, 512)[:, rows_].sum(0)

    class Tower:
        def __init__(self, scale):
            self.w, self.scale = torch.randn(3 * 8 * 8, 16, generator=gen, dtype=torch.float64), scale

        def encode_image_fwd(self, u, save):
            return u.flatten(1) @ self.w, (u.shape if save else None)

        def encode_image_bwd(self, saved, d):
            return (d @ self.w.t()).reshape(saved)

    class Loss:
        def __init__(self, scale):
            self.scale = scale

        def loss_and_grad(self, e_s, e_t, c, inv_count):
            e_t = e_t.detach().requires_grad_(True)
            part = -c * inv_count * torch.nn.functional.cosine_similarity(e_t - e_s, text).sum()
            d, = torch.autograd.grad(part, e_t)
            return part.detach().reshape(1), d * self.scale, torch.tensor([self.scale], dtype=torch.float64)

    monkeypatch.setattr(direction.resample, 'unprocess_fwd', lambda img, mode='unprocess': 2.0 * img)
    monkeypatch.setattr(direction.resample, 'unprocess_bwd', lambda g, img, unscale=None, mode='unprocess': 2.0 * g / unscale)
    f = object.__new__(direction.DirectionFinder)
    f.device, f.engine, f.until_k, f.noise_mode, f.micro_batch, f.rows = torch.device('cpu'), Engine(), 3, 'const', 2, rows
    f.clip_loss_coef, f.overlap, f._side = coef, overlap, None
    f.clips = [(Tower(s), Loss(s), w) for s, w in zip((64.0, 4096.0), direction.DOUBLE_CLIP_WEIGHTS)][:towers]
    f.delta = 0.1 * torch.randn(1, 8, 512, generator=gen)
    styles = torch.randn(n, 26, 512, generator=gen)
    grad, part = f.loss_and_grad(styles, global_count=n)

    delta = f.delta.double().clone().requires_grad_(True)
    d = torch.zeros(1, 26, 512, dtype=torch.float64).index_put((torch.tensor([0]).view(1, 1), torch.tensor(rows).view(1, -1)), delta)
    img = lambda s: 2.0 * (s.reshape(n, -1) @ A)
    want = 0
    for tower, _, w in f.clips:
        e = img(styles.double() + d) @ tower.w - img(styles.double()) @ tower.w
        want = want - w * coef / n * torch.nn.functional.cosine_similarity(e, text).sum()
    g_want, = torch.autograd.grad(want, delta)
    assert abs(part.item() - want.item()) <= 1e-5 * abs(want.item())      # part_sum is a float32 accumulator
    assert ((grad.double() - g_want[0]).norm() / g_want.norm()).item() <= 1e-5


@pytest.mark.parametrize('overlap', [False, True])
def test_direction_finder_source_key_skips_the_original_branch(overlap, monkeypatch):
    """Host logic of ``loss_and_grad(source_key=...)`` with stand-in engines on the CPU: the original-image branch (find_direction.py:311-312) runs
    once per (key, micro-batch), later calls reuse its embeddings and return the same loss and gradient; without a key nothing is cached."""
    import contextlib
    from stylemc_b200 import direction

    class FakeStream:
        def wait_stream(self, other):
            pass
    if overlap:
        monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: FakeStream())
        monkeypatch.setattr(torch.cuda, 'Stream', lambda device=None: FakeStream())
        monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
        monkeypatch.setattr(torch.Tensor, 'record_stream', lambda self, s: None)
    gen = torch.Generator().manual_seed(3)
    A = torch.randn(26 * 512, 3 * 4 * 4, generator=gen, dtype=torch.float64) * 0.02
    W = torch.randn(3 * 4 * 4, 16, generator=gen, dtype=torch.float64)
    text = torch.nn.functional.normalize(torch.randn(1, 16, generator=gen, dtype=torch.float64), dim=1)
    calls = dict(nograd=0, grad=0)

    class Engine:
        def forward(self, s, until_k, noise_mode, save, grad_rows=None):
            calls['grad' if save else 'nograd'] += 1
            return None, (s.double().reshape(len(s), -1) @ A).reshape(len(s), 3, 4, 4), s

        def backward(self, saved, g_img, rows_, noise_mode):
            return (g_img.reshape(len(g_img), -1) @ A.t()).reshape(len(g_img), 26, 512)[:, rows_].sum(0)

    class Tower:
        def encode_image_fwd(self, u, save):
            return u.flatten(1) @ W, (u.shape if save else None)

        def encode_image_bwd(self, saved, d):
            return (d @ W.t()).reshape(saved)

    class Loss:
        def loss_and_grad(self, e_s, e_t, c, inv_count):
            e_t = e_t.detach().requires_grad_(True)
            part = -c * inv_count * torch.nn.functional.cosine_similarity(e_t - e_s, text).sum()
            d, = torch.autograd.grad(part, e_t)
            return part.detach().reshape(1), d, torch.ones(1, dtype=torch.float64)

    monkeypatch.setattr(direction.resample, 'unprocess_fwd', lambda img, mode='unprocess': img)
    monkeypatch.setattr(direction.resample, 'unprocess_bwd', lambda g, img, unscale=None, mode='unprocess': g / unscale)
    f = object.__new__(direction.DirectionFinder)
    f.device, f.engine, f.until_k, f.noise_mode, f.micro_batch, f.rows = torch.device('cpu'), Engine(), 3, 'const', 2, direction.S_TRAINABLE_SPACE_CHANNELS
    f.clip_loss_coef, f.overlap, f._side, f._src_cache = 1.0, overlap, None, {}
    f.clips = [(Tower(), Loss(), 1.0)]
    f.delta = 0.1 * torch.randn(1, 8, 512, generator=gen)
    styles = torch.randn(5, 26, 512, generator=gen)                      # three micro-batches (2 + 2 + 1)
    g0, p0 = f.loss_and_grad(styles, global_count=5)
    assert calls == dict(nograd=3, grad=3) and not f._src_cache
    g1, p1 = f.loss_and_grad(styles, global_count=5, source_key='batch 7')
    assert calls == dict(nograd=6, grad=6) and len(f._src_cache) == 3
    g2, p2 = f.loss_and_grad(styles, global_count=5, source_key='batch 7')
    assert calls == dict(nograd=6, grad=9)                               # the original-image branch did not run again
    for g, p_ in ((g1, p1), (g2, p2)):
        assert torch.equal(g, g0) and torch.equal(p_, p0)
    f.loss_and_grad(styles[:3], global_count=3, source_key='batch 8')      # another key: computed (2 micro-batches) and kept
    assert calls['nograd'] == 8 and len(f._src_cache) == 5
    f.clear_source_cache()
    assert not f._src_cache


def test_npz_formats_and_find_direction_loop(tmp_path):
    """the w / s / direction npz layouts (generate_w.py:51, w_s_converter.py:82, find_direction.py:260,334,351,
    generate_fromS.py:114,125) and the optimisation loop's bookkeeping (iteration count, cosine LR, batch slices, checkpoints,
    final file, resume) with a stand-in step function."""
    import numpy as np
    from stylemc_b200 import direction, npzio as io
    gen = torch.Generator().manual_seed(1)
    S = torch.randn(11, 26, 512, generator=gen)
    io.save_styles(tmp_path / 's.npz', S)
    assert list(np.load(tmp_path / 's.npz').keys()) == ['s'] and np.load(tmp_path / 's.npz')['s'].dtype == np.float32
    assert torch.equal(io.load_styles(tmp_path / 's.npz'), S) and io.load_styles(tmp_path / 's.npz', n=4).shape[0] == 4
    ws = torch.randn(3, 18, 512, generator=gen)
    io.save_w(tmp_path / 'w.npz', ws)
    assert torch.equal(io.load_w(tmp_path / 'w.npz'), ws)
    with pytest.raises(RuntimeError):
        io.save_direction(tmp_path / 'bad.npz', torch.zeros(2, 26, 512))
    np.savez(tmp_path / 'bad_s.npz', s=np.zeros((3, 20, 512), np.float32))
    with pytest.raises(RuntimeError):
        io.load_styles(tmp_path / 'bad_s.npz')
    assert io.direction_path('out', 'a happy face') == os.path.join('out', 'direction_a_happy_face.npz')

    f = object.__new__(direction.DirectionFinder)
    f.device, f.rows, f.lr, f.world, f.group = torch.device('cpu'), direction.S_TRAINABLE_SPACE_CHANNELS, 1.5, 1, None
    f.delta = torch.zeros(1, 8, 512)
    calls = []

    def step(styles, lr=None, global_count=None, source_key=None):
        calls.append((styles.shape[0], lr, global_count, styles[0, 0, 0].item()))
        f.delta += 1.0
        return dict(loss=torch.tensor(0.0))
    f.step = step
    out = tmp_path / 'run'
    final = io.find_direction(f, S, batch_size=4, n_epochs=5, outdir=str(out), text_prompt='a happy face', seed=3, checkpoint_every=6, zero_init='keep')
    total = math.ceil(11 / 4) * 5                                                     # find_direction.py:286-287
    assert len(calls) == total
    rng = np.random.RandomState(3)
    for it, (n, lr, count, first) in enumerate(calls, 1):
        i = rng.randint(0, 3)
        assert n == count == min(4, 11 - 4 * i) and first == S[4 * i, 0, 0].item()   # :303-304, ragged last batch
        assert abs(lr - o_dir.cosine_lr(1.5, it, total)) < 1e-12                      # :298-299
    want = torch.zeros(1, 26, 512)
    want[:, f.rows] = float(total)
    assert torch.equal(final, want) and torch.equal(io.load_direction(io.direction_path(str(out), 'a happy face')), want)
    last = io.load_direction(out / 'direction_last.npz')                              # written at iterations 5 and 11 (it % 6 == 5, as :333)
    assert last[0, f.rows[0], 0].item() == 11.0 and last[0, 0].abs().max().item() == 0.0
    # resume (:266-270): the trainable rows of the saved direction become delta
    f.delta.zero_()
    calls.clear()
    io.find_direction(f, S, batch_size=11, n_epochs=1, resume=str(out / 'direction_last.npz'))
    assert len(calls) == 1 and f.delta[0, 0, 0].item() == 12.0
    # start from delta == 0 (find_direction.py:270): the directional loss is 0/0 there (NaN in the reference, clip_loss.py:27-28; zero
    # gradient in smc_clip_loss), so the loop seeds delta -- or raises / keeps it when asked to
    f.delta.zero_()
    calls.clear()
    with pytest.raises(RuntimeError, match='delta == 0'):
        io.find_direction(f, S, batch_size=11, n_epochs=1, zero_init='raise')
    assert not calls
    seen = []
    f.step = lambda styles, lr=None, global_count=None, source_key=None: seen.append(f.delta.clone()) or dict(loss=torch.tensor(0.0))
    with pytest.warns(UserWarning, match='delta == 0'):
        io.find_direction(f, S, batch_size=11, n_epochs=1)
    assert len(seen) == 1 and seen[0].abs().min().item() > 0 and 0.005 < seen[0].std().item() < 0.02
    g = object.__new__(direction.DirectionFinder)
    g.device, g.delta = torch.device('cpu'), torch.zeros(1, 8, 512)
    g.seed_delta()
    assert torch.equal(g.delta, seen[0])                                              # seeded on the host: the same on every rank
    with pytest.raises(ValueError):
        io.find_direction(f, S, batch_size=11, n_epochs=1, zero_init='maybe')


def test_save_canvases_writes_the_reference_file_names(tmp_path):
    from PIL import Image
    from stylemc_b200 import npzio as io
    canv = torch.randint(0, 256, (2, 16, 32, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    paths = io.save_canvases(canv, str(tmp_path), 'a happy face', first_index=7)
    assert [os.path.basename(p) for p in paths] == ['a_happy_face_007.jpeg', 'a_happy_face_008.jpeg']     # generate_fromS.py:205
    assert Image.open(paths[0]).size == (32, 16)
    with pytest.raises(RuntimeError):
        io.save_canvases(canv.float(), str(tmp_path), 'x')


def test_styles_dict_view_round_trip():
    """BASELINE.json's "per-layer styles dict": a named view of the reference's [N, 26, 512] S tensor (SURVEY.md 8a naming note)."""
    from stylemc_b200 import utils
    G = o_syn.make_generator(32, seed=1, channel_base=1024, channel_max=64)
    ws = torch.randn(3, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(2))
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    d = utils.styles_dict(G, S, shapes)
    assert list(d) == ['b4.conv1', 'b4.torgb', 'b8.conv0', 'b8.conv1', 'b8.torgb', 'b16.conv0', 'b16.conv1', 'b16.torgb',
                       'b32.conv0', 'b32.conv1', 'b32.torgb']
    assert d['b8.conv0'].shape == (3, shapes[1][0]) and d['b8.conv0'].data_ptr() == S[:, 2].data_ptr()       # a view of row 2
    assert d['b4.torgb'].shape[1] == shapes[0][2]
    assert torch.equal(utils.styles_from_dict(G, d, shapes), S)
    with pytest.raises(RuntimeError):
        utils.styles_from_dict(G, {k: v for k, v in d.items() if k != 'b8.torgb'})


def test_plugin_boundary_is_registered_as_torch_library_ops():
    """torch.ops.stylemc_b200.bias_act / .upfirdn2d carry the argument lists of the reference's pybind plugin functions
    (bias_act.cpp:32, upfirdn2d.cpp:16), have fake (meta) implementations for tracing and no CPU kernel."""
    import stylemc_b200.ops  # noqa: F401
    s = str(torch.ops.stylemc_b200.bias_act.default._schema)
    assert [a.split()[-1] for a in s[s.index('(') + 1:s.index(')')].split(', ')] == ['x', 'b', 'xref', 'yref', 'dy', 'grad', 'dim', 'act', 'alpha',
                                                                                   'gain', 'clamp']
    s = str(torch.ops.stylemc_b200.upfirdn2d.default._schema)
    assert [a.split()[-1] for a in s[s.index('(') + 1:s.index(')')].split(', ')] == ['x', 'f', 'upx', 'upy', 'downx', 'downy', 'padx0', 'padx1', 'pady0',
                                                                                   'pady1', 'flip', 'gain']
    x, f = torch.empty(2, 3, 8, 8, device='meta'), torch.empty(4, 4, device='meta')
    assert torch.ops.stylemc_b200.upfirdn2d(x, f, 2, 2, 1, 1, 2, 1, 2, 1, False, 4.0).shape == (2, 3, 16, 16)        # upsample2d
    assert torch.ops.stylemc_b200.upfirdn2d(x, f, 1, 1, 2, 2, 1, 1, 1, 1, False, 1.0).shape == (2, 3, 4, 4)          # downsample2d
    xl = torch.empty(2, 8, 9, 9, device='meta').to(memory_format=torch.channels_last)
    y = torch.ops.stylemc_b200.upfirdn2d(xl, f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0)                                  # the conv0 FIR: (2H+1)^2 -> (2H)^2
    assert y.shape == (2, 8, 8, 8) and y.stride(1) == 1
    assert torch.ops.stylemc_b200.bias_act(x, None, None, None, None, 0, 1, 3, 0.2, 2 ** 0.5, 256.0).shape == x.shape
    with pytest.raises(NotImplementedError):                                     # no CPU kernel, no fallback
        torch.ops.stylemc_b200.bias_act(torch.zeros(2, 3), None, None, None, None, 0, 1, 3, 0.2, 1.0, -1.0)


def test_mapping_network_and_generate_w():
    """SURVEY.md 8(f) row 1, z -> W+ (generate_w.py:46-51; upstream MappingNetwork with legacy.py:129-136 kwargs): the oracle against an
    independent float64 evaluation, the invariances of the architecture, parameter names of legacy.py:175-181, and identical weights in
    the product's module (whose forward needs the bias_act kernel: tests/test_zz_clip_b16_gpu.py)."""
    import numpy as np
    from stylemc_b200 import networks
    G = o_syn.make_generator(32, seed=1, channel_base=1024, channel_max=64, mapping=True)
    plain = o_syn.make_generator(32, seed=1, channel_base=1024, channel_max=64)
    assert all(torch.equal(v, G.state_dict()[k]) for k, v in plain.state_dict().items())       # the mapping network does not move the synthesis weights
    names = [k for k in G.state_dict() if k.startswith('mapping.')]
    assert names == ['mapping.w_avg'] + [f'mapping.fc{i}.{p}' for i in range(8) for p in ('weight', 'bias')]
    seeds = [1, 2, 5]
    ws = o_syn.generate_w(G, seeds, truncation_psi=0.7)
    assert ws.shape == (3, G.synthesis.num_ws, 512) and all(torch.equal(ws[:, 0], ws[:, j]) for j in range(ws.shape[1]))
    # float64, written out: z / rms(z) -> 8 x lrelu(x W^T * 0.01 / sqrt(512) + 0.01 b) * sqrt(2) -> w_avg + psi (w - w_avg)
    x = np.concatenate([np.random.RandomState(s).randn(1, 512) for s in seeds])
    x = x / np.sqrt((x * x).mean(1, keepdims=True) + 1e-8)
    for i in range(8):
        fc = getattr(G.mapping, f'fc{i}')
        x = x @ (fc.weight.double().numpy().T * (0.01 / math.sqrt(512))) + fc.bias.double().numpy() * 0.01
        x = np.where(x > 0, x, 0.2 * x) * math.sqrt(2)
    avg = G.mapping.w_avg.double().numpy()
    want = avg + 0.7 * (x - avg)
    assert np.abs(ws[:, 0].double().numpy() - want).max() <= 1e-5
    z = torch.randn(2, 512, generator=torch.Generator().manual_seed(3))
    assert torch.allclose(G.mapping(z), G.mapping(3.0 * z), atol=1e-5)                        # second-moment normalisation
    cut = G.mapping(z, truncation_psi=0.5, truncation_cutoff=2)
    full = G.mapping(z)
    assert torch.equal(cut[:, 2:], full[:, 2:]) and torch.allclose(cut[:, :2], G.mapping.w_avg.lerp(full[:, :2], 0.5))
    Gp = networks.make_generator(32, seed=1, channel_base=1024, channel_max=64, mapping=True)
    assert all(torch.equal(v, Gp.state_dict()[k]) for k, v in G.state_dict().items())
