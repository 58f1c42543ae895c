"""CPU: the launch planner of the halo-tile convolution kernel (csrc/hconv.cu, reached through the host-only entry ``smc_igemm_plan``)
on every GEMM shape of one find_direction step of the 1024-px config-f network and of the CLIP ViT-B/32 towers: a plan exists, fits the
227 KB of shared memory and the 512 TMEM columns, and the number of accumulator commits the MMA issuer makes per tile equals the number
of drains the epilogue warps wait for (a mismatch deadlocks the kernel).  No GPU work: the planner is plain host code."""
import ctypes

import pytest

from stylemc_b200 import _lib, gemm

CH = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}   # config-f: min(32768 // res, 512)
EUNSUPPORTED = -2


def make_desc(n, H, W, C, n_out, taps, x3=True, acc=512, problems=None, a_imgs=None, HA=None, WA=None, rows_b=None, mask=False):
    d = _lib.IgemmDesc()
    planes = 2 if x3 else 1
    a_imgs = a_imgs or n
    d.A, d.NA, d.HA, d.WA, d.C, d.lda = 0x10000, a_imgs * planes, HA or H, WA or W, C, C
    rows_b = rows_b or 9 * n_out
    d.B, d.rowsB, d.ldb = 0x20000, rows_b * planes, C
    d.n_img, d.H, d.W, d.n_out = n, H, W, n_out
    full = [(t if len(t) == 4 else (0,) + tuple(t)) for t in taps]
    full = [(dn, dy, dx, ti * n_out) for dn, dy, dx, ti in full]
    if x3:
        full = full + [(dn, dy, dx, br + rows_b) for dn, dy, dx, br in full] + [(dn + a_imgs, dy, dx, br) for dn, dy, dx, br in full]
    d.ntaps = len(full)
    for i, (dn, dy, dx, br) in enumerate(full):
        d.taps[i].dn, d.taps[i].dy, d.taps[i].dx, d.taps[i].brow = dn, dy, dx, br
    e = d.epi
    e.o_sn, e.o_sh, e.o_sw = H * W * n_out, W * n_out, n_out
    if mask:                 # fused activation backward of the dgrad GEMMs
        e.mask_y, e.post_scale, e.out_hi, e.out_lo = 0x60000, 0x70000, 0x80000, 0x90000
    else:
        e.out_f32 = 0x40000
    d.acc_chunk_k = acc
    if problems:
        d.nprob = len(problems)
        for q, (nt, off) in enumerate(problems):
            d.prob_ntaps[q], d.prob_o_off[q] = nt, off
    return d


def plan(d):
    o = _lib.IgemmPlanInfo()
    rc = _lib.lib().smc_igemm_plan(ctypes.byref(d), ctypes.byref(o))
    return rc, o


def parity_group(n_out_plane_elems=4096):
    taps, problems = [], []
    for q, (r, c) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        t = gemm.up2_parity_taps(r, c)
        taps += t
        problems.append((len(t), q * n_out_plane_elems))
    return taps, problems


def check(o, base_taps, x3, halo):
    assert o.kernel == 1
    assert o.smem_bytes <= 227 * 1024
    assert 32 <= o.tmem_cols <= 512 and o.tmem_cols & (o.tmem_cols - 1) == 0
    if o.pair:       # CTA-pair launch (cta_group::2): clusters of two CTAs over pairs of images, 128-wide N tiles, never the merged-B mode
        assert o.bn == 128 and o.mode != 2 and not o.b_resident and not o.a_share
        assert 2 <= o.grid <= 148 and o.grid % 2 == 0 and o.grid // 2 <= o.super_tiles
    else:
        assert 1 <= o.grid <= 148 and o.grid <= o.super_tiles
    assert o.Wp - o.Wt == halo and o.Wt <= 80          # 64, or ceil(W / floor(W / 64)) for the 2^k + 1 wide parity planes (no 1-pixel column tile)
    assert o.na_hi >= 2 and (o.na_lo >= 1 if x3 else o.na_lo == 0)
    assert sum(o.prob_ntaps[q] for q in range(o.nprob)) == base_taps
    for q in range(o.nprob):
        assert o.prob_ndrains[q] >= 1
        assert o.prob_commits[q] == o.prob_ndrains[q], (q, list(o.prob_commits), list(o.prob_ndrains))
    if o.b_resident:
        passes = 2 if o.mode == 1 else 1
        assert o.nb == o.kchunks * base_taps * passes <= 32
        assert [o.prob_stages[q] for q in range(o.nprob)] == [o.kchunks * passes * sum(o.prob_ntaps[j] for j in range(q)) for q in range(o.nprob)]
    if o.a_share:
        assert o.nprob > 1 and o.kchunks == 1 and o.mode != 1


@pytest.mark.parametrize('n', [1, 16, 64])
@pytest.mark.parametrize('x3', [True, False])
def test_every_synthesis_gemm_of_the_1024_network_has_a_plan(n, x3):
    for res in (4, 8, 16, 32, 64, 128, 256, 512, 1024):
        c = CH[res]
        acc = 0 if not x3 else (64 if res <= 16 else 512)
        # conv1 forward and its dgrad (with the fused activation backward where the engine uses it)
        for taps, mask in ((gemm.TAPS_3X3, False), (gemm.TAPS_3X3_DGRAD, False), (gemm.TAPS_3X3_DGRAD, True)):
            rc, o = plan(make_desc(n, res, res, c, c, taps, x3, acc if taps is gemm.TAPS_3X3 else (512 if x3 else 0), mask=mask))
            if res == 4:
                assert rc == EUNSUPPORTED and o.kernel == 0          # 4 x 4 planes stay on the per-tap kernel (csrc/igemm.cu)
                continue
            assert rc == 0, (res, rc)
            check(o, 9, x3, halo=2)
            assert bool(o.pair) == (c >= 128 and n % 2 == 0)         # image pairs: an odd batch keeps the single-CTA kernel
            assert (o.mode == 0) == (not x3)
            if x3:
                assert o.mode == (2 if c <= 64 else 1)                  # merged-B below 128 output channels
        if res == 4:
            continue
        # conv0: the four output parities of the stride-2 transposed conv as one problem group
        hin, cin = res // 2, CH[res // 2]
        taps, problems = parity_group((hin + 1) * (hin + 1) * c * n)
        rc, o = plan(make_desc(n, hin + 1, hin + 1, cin, c, taps, x3, acc, problems=problems, HA=hin, WA=hin))
        if hin + 1 < 8:
            assert rc == EUNSUPPORTED
        else:
            assert rc == 0, (res, rc)
            check(o, 9, x3, halo=1)
            assert o.nprob == 4 and [o.prob_ntaps[q] for q in range(4)] == [4, 2, 2, 1]
            if hin + 1 > 64:            # 65 / 129 / 257 / 513 columns: 1 / 2 / 4 / 8 column tiles of 65, not 64-wide tiles plus a 1-pixel one
                assert -(-(hin + 1) // o.Wt) == (hin + 1) // 64, (hin + 1, o.Wt)
            assert bool(o.a_share) == (cin == 64 and (c <= 64 or not x3))
        # ... and its dgrad from the four gradient parity planes (four A sources)
        rc, o = plan(make_desc(n, hin, hin, c, cin, gemm.up2_dgrad_taps(n), x3, 512 if x3 else 0, a_imgs=4 * n, HA=hin + 1, WA=hin + 1, rows_b=9 * cin))
        if hin < 8:
            assert rc == EUNSUPPORTED
        else:
            assert rc == 0, (res, rc)
            check(o, 9, x3, halo=1)


@pytest.mark.parametrize('rows', [64 * 50, 2 * 64 * 50, 64 * 49, 64 * 197, 64 * 196, 3 * 197])      # ViT-B/32 and ViT-B/16 token / patch rows
def test_clip_linears_have_a_plan(rows):
    for k, n_out in ((768, 2304), (768, 768), (768, 3072), (3072, 768)):
        rc, o = plan(make_desc(1, 1, rows, k, n_out, gemm.TAPS_1X1, True, 512, rows_b=n_out))
        if rows >= 1024 and any(rows % c == 0 for c in range(16, 65)):
            assert rc == 0, (rows, k, n_out, rc)
            check(o, 1, True, halo=0)
            assert o.mode == 1 and o.bn == 128
        else:
            assert rc in (0, EUNSUPPORTED)


def test_bad_problem_tables_are_refused():
    taps, problems = parity_group()
    d = make_desc(2, 33, 33, 64, 32, taps, True, 512, problems=problems, HA=32, WA=32)
    d.prob_ntaps[3] = 2                                   # 4 + 2 + 2 + 2 != 9 base taps
    assert plan(d)[0] == -1
    d.prob_ntaps[3] = 1
    d.nprob = 5
    assert plan(d)[0] == -1
