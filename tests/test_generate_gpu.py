"""GPU parity of the batched generate_fromS render (stylemc_b200.generate) against the reference's per-style loop restated on the CPU
oracle (generate_fromS.py:147-175,206): uint8 canvases original | edited.  Integer output: identical except where the fp32 image sits
within rounding distance of an integer boundary (the images agree to ~1e-5), so at most a 1-level difference on a tiny fraction."""
import pytest
import torch

from oracle import synthesis as o_syn

pytestmark = pytest.mark.gpu


def test_generate_fromS_uint8_canvas():
    from stylemc_b200 import generate
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    ws = torch.randn(5, G.synthesis.num_ws, 512, generator=torch.Generator().manual_seed(31))
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    direction = torch.zeros(1, 26, 512)
    direction[:, [2, 3, 5, 6, 8, 9, 11, 12]] = 0.3 * torch.randn(1, 8, 512, generator=torch.Generator().manual_seed(32))
    power = 2.5
    out = generate.generate_fromS(G, S, direction, power, batch=2).cpu()      # ragged last batch on purpose
    assert out.shape == (5, 64, 128, 3) and out.dtype == torch.uint8
    want = []
    for i in range(5):                                                              # the reference's loop: one style at a time
        halves = []
        for g in (0.0, power):
            _, img = o_syn.generate_image(G, 100, S[[i]] + direction * g, shapes, 'const')
            halves.append((img.permute(0, 2, 3, 1) * 127.5 + 128).clamp(0, 255)[0].to(torch.uint8))
        want.append(torch.cat(halves, dim=1))
    want = torch.stack(want)
    diff = (out.int() - want.int()).abs()
    frac = (diff > 0).float().mean().item()
    print(f'uint8 canvas: {frac:.2e} of the values differ, max diff {diff.max().item()}')
    assert diff.max().item() <= 1 and frac <= 2e-3
    assert (out[:, :, :64] != out[:, :, 64:]).any()                               # the edit changed the image


def test_generate_fromS_1024_canvas_golden(golden):
    """BASELINE.json configs[1] network (1024 px) against the reference's own render loop (generate_fromS.py:147-175,206 driving the
    real utils.generate_image, tests/golden/config4.npz): a 16-strided sample of the uint8 canvas, a full-resolution window across the
    original | edited seam and the per-channel sums of the whole canvas."""
    from stylemc_b200 import generate
    g = golden('config4')
    G = o_syn.make_generator(1024, seed=0)
    ws = torch.as_tensor(g['ws'])[:2]
    S, shapes = o_syn.get_styles(G, ws, o_syn.split_ws(G, ws))
    direction = torch.zeros(1, 26, 512)
    direction[:, [2, 3, 5, 6, 8, 9, 11, 12]] = torch.as_tensor(g['delta'])
    out = generate.generate_fromS(G, S, direction, float(g['canvas_power']), batch=2).cpu()
    assert out.shape == (2, 1024, 2048, 3) and out.dtype == torch.uint8
    for tag, got in (('canvas_strided', out[:, ::16, ::16]), ('canvas_window', out[:, 448:576, 896:1152])):
        diff = (got.int() - torch.as_tensor(g[tag]).int()).abs()
        frac = (diff > 0).float().mean().item()
        print(f'{tag}: {frac:.2e} of the values differ, max diff {diff.max().item()}')
        assert diff.max().item() <= 1 and frac <= 2e-3
    sums = out.long().sum(dim=(1, 2))
    want = torch.as_tensor(g['canvas_sum'])
    assert ((sums - want).abs().float() / want.float()).max().item() <= 1e-5      # +-1 level on <= 0.2 % of 2M values per channel


def test_generate_fromS_argument_errors():
    from stylemc_b200 import generate
    G = o_syn.make_generator(64, seed=1, channel_base=2048, channel_max=512)
    with pytest.raises(RuntimeError):
        generate.generate_fromS(G, torch.zeros(2, 26, 512), torch.zeros(2, 26, 512), 1.0)
