"""GPU parity of the fused synthesis engine against the reference's own output (tests/golden/synth64.npz: the reference's
utils.generate_image driving the restated network modules on CPU, impl='ref' ops) and against the CPU oracle's autograd."""
import numpy as np
import pytest
import torch

from oracle import synthesis as o_syn

pytestmark = pytest.mark.gpu

TRAINABLE = [2, 3, 5, 6, 8, 9, 11, 12]     # find_direction.py:41
IMG_TOL = {'x1': 1e-2, 'mixed': 1e-2, 'x3': 2e-4, 'x3p': 2e-5}   # north_star: images <= 1e-2 max-abs for 16-bit operand paths


def small_net():
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)      # same as oracle/pin_reference.make_small_net
    shapes = o_syn.get_temp_shapes(G)
    return G, shapes


@pytest.mark.parametrize('precision', ['x1', 'mixed', 'x3', 'x3p'])
def test_generate_image_golden(golden, precision):
    from stylemc_b200 import utils
    g = golden('synth64')
    G, shapes = small_net()
    assert [tuple(s) for s in g['temp_shapes']] == shapes
    styles = torch.as_tensor(g['styles']).cuda()
    xs, img = utils.generate_image(G, 100, styles, shapes, 'const', 'cuda', precision=precision)
    ref = torch.as_tensor(g['img'])
    err = (img.cpu() - ref).abs().max().item()
    print(f'{precision}: img max-abs err {err:.3e} (ref range {ref.min():.2f}..{ref.max():.2f})')
    assert img.shape == ref.shape and err <= IMG_TOL[precision]
    for i, x in enumerate(xs):
        r = torch.as_tensor(g[f'xs{i}'])
        e = (x.cpu() - r).abs().max().item() / r.abs().max().item()
        print(f'  xs[{i}] rel err {e:.3e}')
        assert e <= {'x1': 5e-3, 'mixed': 5e-3, 'x3': 1e-4, 'x3p': 1e-5}[precision]
    _, img2 = utils.generate_image(G, 2, styles, shapes, 'const', 'cuda', precision=precision)
    assert (img2.cpu() - torch.as_tensor(g['img_k2'])).abs().max().item() <= IMG_TOL[precision]


@pytest.mark.parametrize('precision,tol', [('x1', 5e-2), ('x3', 5e-3), ('x3p', 1e-3)])
def test_style_gradient_vs_oracle_autograd(golden, precision, tol):
    """d(sum(img * g)) / d(delta) for delta added to the trainable S rows, batch-summed (find_direction.py:307-308,336).

    The gradient is far more sensitive than the image: a forward rounding error eps flips the lrelu slope of a fraction ~eps of
    the units, which perturbs the gradient by ~sqrt(eps) in relative L2.  fp16 operands (eps ~3e-4) therefore give ~1e-2,
    split operands (~2e-5, limited by the tensor core's truncating accumulation) ~3e-3, and split operands with promoted
    accumulation (~7e-7) reach the 1e-3 of BASELINE.json; the fp32 reference itself is 1.7e-4 from the fp64 truth."""
    from stylemc_b200 import utils
    g = golden('synth64')
    G, shapes = small_net()
    styles = torch.as_tensor(g['styles'])
    gen = torch.Generator().manual_seed(11)
    delta = (0.1 * torch.randn(1, 8, 512, generator=gen)).requires_grad_(True)
    g_img = torch.randn(styles.shape[0], 3, 64, 64, generator=gen)
    direction = torch.zeros(1, 26, 512).index_put((torch.tensor([0]).view(1, 1), torch.tensor(TRAINABLE).view(1, -1)), delta)
    styles2 = styles + direction
    _, img_ref = o_syn.generate_image(G, 100, styles2, shapes, 'const')
    (img_ref * g_img).sum().backward()
    ref = delta.grad[0]

    eng = utils.engine_for(G, 'cuda', precision)
    _, img, saved = eng.forward(styles2.detach().cuda(), save=True)
    grad = eng.backward(saved, g_img.cuda(), TRAINABLE).cpu()
    assert (img.cpu() - img_ref.detach()).abs().max().item() <= IMG_TOL[precision]
    for i, row in enumerate(TRAINABLE):
        e = ((grad[i] - ref[i]).norm() / ref[i].norm()).item()
        print(f'{precision}: row {row} grad rel-l2 err {e:.3e}  |ref| {ref[i].norm():.3e}')
    rel = ((grad - ref).norm() / ref.norm()).item()
    print(f'{precision}: total grad rel-l2 err {rel:.3e}')
    assert rel <= tol


def test_fused_torgb_over_several_n_tiles(monkeypatch):
    """conv1 layers of 256 / 512 channels (2 / 4 N tiles of the halo-tile kernel) take ToRGB in their epilogue with one partial-sum image per
    N tile (smc_igemm_epilogue::rgb_snt, summed in index order by smc_img_finish): bit-identical from run to run, and equal to the separate
    smc_torgb pass (STYLEMC_FUSE_RGB_WIDE=0) to fp32 rounding, forward and backward (saved ToRGB clamp mask instead of the recomputation)."""
    from stylemc_b200 import synthesis
    G = o_syn.make_generator(64, seed=3, channel_base=16384, channel_max=512)      # 32 px: 512 channels, 64 px: 256 channels
    ws = torch.randn(3, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(5))
    S, _ = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    S = S.cuda()
    g_img = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(6)).cuda()

    def run(wide):
        monkeypatch.setenv('STYLEMC_FUSE_RGB_WIDE', '1' if wide else '0')
        eng = synthesis.SynthesisEngine(G, 'cuda', precision='x3p')
        assert eng.fuse_rgb_wide == wide
        _, img, saved = eng.forward(S, save=True, grad_rows=TRAINABLE)
        assert (len(saved.rgb_pass) > 0) == wide        # 32 / 64 px blocks: fused (with a saved clamp mask) only in the wide mode
        return img, eng.backward(saved, g_img, TRAINABLE)

    img_a, grad_a = run(True)
    img_b, grad_b = run(True)
    assert torch.equal(img_a, img_b)                                      # (the style-gradient sums use float atomics: equal to rounding only)
    assert ((grad_a - grad_b).norm() / grad_b.norm()).item() <= 1e-5
    img_c, grad_c = run(False)
    err = (img_a - img_c).abs().max().item()
    rel = ((grad_a - grad_c).norm() / grad_c.norm()).item()
    print(f'fused vs separate ToRGB: img max-abs diff {err:.2e}, style-gradient rel-l2 {rel:.2e}')
    assert err <= 2e-5 and rel <= 1e-4
