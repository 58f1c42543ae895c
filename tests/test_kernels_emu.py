"""CPU: the tiled attention kernels for long sequences (ViT-B/16, 197 tokens; csrc/vit.cu attention_fwd_rows_kernel,
attention_bwd_q_kernel, attention_bwd_kv_kernel) executed UNCHANGED on host threads by a small CUDA execution-model shim
(tests/emu/cuda_emu.h) and compared with a float64 softmax attention and its gradient.  The kernel text is cut out of vit.cu at
test time, so this checks the index arithmetic and data flow of the code that ships, without a GPU; the whole-sequence kernels
that the GPU tests already verify run through the same shim as a check of the shim itself.  The same binary is also built with
AddressSanitizer + UBSan (out-of-bounds global / shared-memory accesses: memcheck's job on a device) and with ThreadSanitizer (a missing
``__syncthreads`` is a data race between host threads: racecheck's job)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, 'tests', 'emu')
KERNELS = ['attention_fwd2_kernel', 'attention_bwd2_kernel', 'attention_fwd_rows_kernel', 'attention_bwd_q_kernel', 'attention_bwd_kv_kernel']


def _cut(text, start_pattern):
    """The top-level definition that starts at ``start_pattern`` and ends at the first '}' in column 0."""
    m = re.search(start_pattern, text, re.M)
    assert m, start_pattern
    end = text.index('\n}\n', m.start()) + 3
    return text[m.start():end]


GLUE_KERNELS = ['layernorm_fwd_kernel', 'layernorm_bwd_kernel', 'quickgelu_fwd_kernel', 'quickgelu_bwd_kernel', 'split_rows_kernel', 'head_proj_kernel',
                'head_proj_bwd_kernel', 'clip_loss_kernel']


SYNTH_KERNELS = ['demod_kernel', 'pack_nhwc_kernel', ('unpack_nchw_kernel', 'template <class T>'), 'sgrad_sample_kernel', 'sgrad_sum_kernel']


def extract(cu, common_cuh, kernels=KERNELS, host_functions=('attention_block_rows',)):
    parts = ['namespace smc {', _cut(common_cuh, r'^__device__ __forceinline__ float warp_sum\(')]
    if '__forceinline__ void store_split(' in cu:
        parts.append(_cut(cu, r'^__device__ __forceinline__ void store_split\('))
    for k in kernels:
        k, prefix = k if isinstance(k, tuple) else (k, '')
        if k.startswith('@'):                      # a __device__ helper the kernels call
            parts.append(_cut(cu, r'^__device__ __forceinline__ \w+ ' + k[1:] + r'\('))
            continue
        body = _cut(cu, r'^__global__ void __launch_bounds__\([^)]*\) ' + k + r'\(')
        body, n = re.subn(r'extern __shared__ float (\w+)\[\];', r'float* \1 = emu_smem;', body)
        assert n <= 1
        parts.append(prefix + '\n' + body)
    parts.append('}  // namespace smc')
    for f in host_functions:
        parts.append(_cut(cu, r'^static int ' + f + r'\('))
    return '\n'.join(parts) + '\n'


SANITIZERS = {'plain': ['-O2'],
              # pointer-overflow is off: kernels form `nullptr + offset` for an absent lo plane and never dereference it (fine on a device)
              # heap bounds of every global AND shared-memory access (the shim's shared memory is a heap block of exactly the launcher's size)
              'address': ['-O2', '-g', '-fsanitize=address,undefined', '-fno-sanitize=pointer-overflow', '-fno-sanitize-recover=undefined'],
              # the shim's barriers are the only synchronisation, so a missing __syncthreads shows up as a data race (checked by removing one)
              'thread': ['-O2', '-g', '-fsanitize=thread']}


def build_and_run(tmp_path, sanitizer, main_cpp, kernels, host_functions, source='vit.cu', argv=()):
    csrc = os.path.join(ROOT, 'stylemc_b200', 'csrc')
    inc = extract(open(os.path.join(csrc, source)).read(), open(os.path.join(csrc, 'common.cuh')).read(), kernels, host_functions)
    (tmp_path / 'kernels_extracted.inc').write_text(inc)
    exe = str(tmp_path / 'emu')
    # -fno-strict-aliasing: CUDA code type-puns vector registers (uint4 <-> __half2[4]); g++ -O2 miscompiles that under its aliasing rules
    cc = subprocess.run(['g++', '-std=c++20', '-pthread', '-Wno-unknown-pragmas', '-fno-strict-aliasing'] + SANITIZERS[sanitizer] +
                        ['-I', str(tmp_path), '-I', EMU, os.path.join(EMU, main_cpp), '-o', exe], capture_output=True, text=True)
    if cc.returncode != 0 and sanitizer != 'plain':
        pytest.skip(f'-fsanitize={sanitizer} runtime not available: {cc.stderr[-200:]}')
    assert cc.returncode == 0, cc.stderr
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66', ASAN_OPTIONS='detect_leaks=0')
    r = subprocess.run([exe] + list(argv), capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout, r.stderr[-3000:])
    assert r.returncode == 0, r.stdout + r.stderr[-3000:]
    assert 'FAIL' not in r.stdout
    assert 'ThreadSanitizer' not in r.stderr and 'AddressSanitizer' not in r.stderr and 'runtime error' not in r.stderr
    return r.stdout


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', list(SANITIZERS))
def test_tiled_attention_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    out = build_and_run(tmp_path, sanitizer, 'attention_main.cpp', KERNELS, ('attention_block_rows',))
    assert out.count('ok  ') == 5


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', list(SANITIZERS))
def test_vit_glue_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    """LayerNorm forward / backward, the embedding head and its transpose, QuickGELU, split_rows and the directional CLIP loss kernel (the
    one with the most barriers in vit.cu) of the measured ViT-B/32 path, same shim, same sanitizers."""
    out = build_and_run(tmp_path, sanitizer, 'vit_glue_main.cpp', GLUE_KERNELS, ())
    assert out.count('ok  ') == 18


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', list(SANITIZERS))
def test_synthesis_glue_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    """Shared-memory glue kernels of csrc/synth.cu on the measured path: demodulation coefficients, the final style-gradient assembly
    (SURVEY.md 8a algebra) and the NCHW <-> NHWC tile transposes, launched with the grids the C layer uses."""
    out = build_and_run(tmp_path, sanitizer, 'synth_glue_main.cpp', SYNTH_KERNELS, (), source='synth.cu')
    assert out.count('ok  ') == 7


FIR_KERNELS = ['@split4', ('fir_act3_kernel', 'template <int C, int JT, int SAVE, bool NOISE, int MINB>')]


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', ['plain', 'address'])
def test_fir_act3_kernel_on_the_cpu_shim(tmp_path, sanitizer):
    """The conv0 tail of the fused synthesis path (csrc/synth.cu fir_act3_kernel: the heaviest HBM-bound kernel of a step) against a
    float64 upfirdn2d + bias_act over the transposed-conv output.  Plane entries outside that output are NaN, the run under
    AddressSanitizer + UBSan checks the bounds and the 8 / 16-byte alignment of every vector access.  (No barriers: no TSan leg.)"""
    out = build_and_run(tmp_path, sanitizer, 'fir_act3_main.cpp', FIR_KERNELS, (), source='synth.cu')
    assert out.count('ok  ') == 3


FIR_BWD_KERNELS = ['@split4', '@ld_h4', ('fir_bwd3_kernel', 'template <int C, int JT, bool LO, int MINB>')]


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', ['plain', 'address'])
def test_fir_bwd3_kernel_on_the_cpu_shim(tmp_path, sanitizer):
    """The transpose of the conv0 FIR (csrc/synth.cu fir_bwd3_kernel) against the float64 adjoint of fir_act3's filter, including the zero
    rows / columns of the planes outside the transposed-conv grid; bounds and vector alignment under AddressSanitizer + UBSan."""
    out = build_and_run(tmp_path, sanitizer, 'fir_bwd3_main.cpp', FIR_BWD_KERNELS, (), source='synth.cu')
    assert out.count('ok  ') == 3


RESAMPLE_KERNELS = ['resample_h_kernel', 'resample_v_kernel', 'resample_vT_kernel', 'resample_hT_kernel', ('resample_rows_kernel', 'template <int OBT, int MAXT>')]


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('in_size,out_size,sanitizer', [(80, 36, 'plain'), (40, 18, 'address'), (20, 36, 'plain')])
def test_unprocess_kernels_on_the_cpu_shim(tmp_path, in_size, out_size, sanitizer):
    """``unprocess`` (find_direction.py:49-52) end to end on the CPU: the shipped resample kernels, launched in the C layer's order on the
    tables stylemc_b200/resample.py builds, against the oracle's F.interpolate(bicubic, antialias) forward and its autograd -- a
    down-sampling case (the benchmark's 1024 -> 224 regime: wide windows) and an up-sampling one (the 64-px tests)."""
    import numpy as np
    import torch
    from oracle import direction as o_dir
    from stylemc_b200 import resample
    gen = torch.Generator().manual_seed(4)
    planes = 3
    x = 0.7 * torch.randn(1, 3, in_size, in_size, generator=gen)                 # ~15 % of the pixels hit the clamp(0, 255)
    g = torch.randn(1, 3, out_size, out_size, generator=gen)
    unscale = 64.0
    start, count, wgt, taps = resample.aa_tables(in_size, out_size)
    oidx, count_t, wgt_t, taps_t = resample.transpose_tables(start, count, wgt, in_size)
    with open(tmp_path / 'in.bin', 'wb') as f:
        f.write(np.array([planes, in_size, out_size, taps, taps_t], np.int32).tobytes())
        for a, dt in ((x.numpy(), np.float32), (start, np.int32), (count, np.int32), (wgt, np.float32), (oidx, np.int32), (count_t, np.int32),
                      (wgt_t, np.float32), ((g * unscale).numpy(), np.float32),
                      (np.array(list(resample.CLIP_MEAN) + list(resample.CLIP_STD) + [unscale]), np.float32)):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
    build_and_run(tmp_path, sanitizer, 'unprocess_main.cpp', RESAMPLE_KERNELS, (), argv=[str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')])
    out = np.fromfile(tmp_path / 'out.bin', np.float32)
    y = torch.from_numpy(out[:planes * out_size * out_size]).reshape(1, 3, out_size, out_size)
    gx = torch.from_numpy(out[planes * out_size * out_size:]).reshape(1, 3, in_size, in_size)
    x64 = x.double().requires_grad_(True)
    y64 = o_dir.unprocess(x64, size=out_size)
    y64.backward(g.double())
    assert (y.double() - y64.detach()).abs().max().item() <= 2e-5
    assert ((gx.double() - x64.grad).norm() / x64.grad.norm()).item() <= 1e-5


ACT_BWD_KERNELS = ['@h8_to_f', '@f_to_h8', '@f_to_h8_split', '@ld8f', ('act_bwd_kernel', 'template <class TG>'), ('act_bwd1_kernel', 'template <class TG>')]


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', list(SANITIZERS))
def test_act_bwd_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    """The activation backward + style-gradient reductions of the fused synthesis path (csrc/synth.cu act_bwd1_kernel and the generic act_bwd_kernel of the 512-channel layers: shared-memory and
    global atomics, partial-warp shuffles) against a float64 restatement, with the launch configuration of smc_act_bwd; ThreadSanitizer
    covers the reductions, AddressSanitizer + UBSan the 16-byte streaming loads and stores."""
    out = build_and_run(tmp_path, sanitizer, 'act_bwd_main.cpp', ACT_BWD_KERNELS, (), source='synth.cu')
    assert out.count('ok  ') == 5


TORGB_KERNELS = ['@h8_to_f', '@f_to_h8', '@f_to_h8_split', '@ld8f', 'torgb_kernel', 'torgb1_kernel']


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', ['plain', 'address'])
def test_torgb_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    """ToRGB + skip-image upsample + next-layer style multiply (csrc/synth.cu torgb1_kernel and the generic torgb_kernel; utils.py:45-49)
    against a float64 restatement, with the launch configurations of smc_torgb."""
    out = build_and_run(tmp_path, sanitizer, 'torgb_main.cpp', TORGB_KERNELS, (), source='synth.cu')
    assert out.count('ok  ') == 5


ACT_BWD_RGB_KERNELS = ['@h8_to_f', '@ld8f', '@split4', ('act_bwd_rgb_kernel', 'template <int C, bool YLO, bool LO>')]


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', ['plain', 'address'])
def test_act_bwd_rgb_kernel_on_the_cpu_shim(tmp_path, sanitizer):
    """The activation backward of the last block (csrc/synth.cu act_bwd_rgb_kernel: at 1024 px the largest activation of a step) against
    a float64 restatement with the launch configuration of launch_act_bwd_rgb.  (No shared memory or shuffles: no TSan leg.)"""
    out = build_and_run(tmp_path, sanitizer, 'act_bwd_rgb_main.cpp', ACT_BWD_RGB_KERNELS, (), source='synth.cu')
    assert out.count('ok  ') == 3


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', ['plain', 'address'])
def test_img_finish_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    """The tail of the fused ToRGB path (csrc/synth.cu img_finish4_kernel / img_finish_kernel: bias, clamp + saved mask, skip-image
    upsample) against a float64 restatement; the kernel is picked as smc_img_finish picks it."""
    out = build_and_run(tmp_path, sanitizer, 'img_finish_main.cpp', ['img_finish_kernel', 'img_finish4_kernel'], (), source='synth.cu')
    assert out.count('ok  ') == 6


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('sanitizer', ['plain', 'address'])
def test_vit_token_kernels_on_the_cpu_shim(tmp_path, sanitizer):
    """patchify / unpatchify for patch sizes 32, 16 (the ViT-B/16 tower) and 8, class-token assembly and the text embedding lookup."""
    out = build_and_run(tmp_path, sanitizer, 'vit_tokens_main.cpp', ['patchify_kernel', 'unpatchify_kernel', 'assemble_tokens_kernel', 'embed_text_kernel'], ())
    assert out.count('ok  ') == 4
