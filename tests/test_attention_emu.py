"""CPU: the tiled attention kernels for long sequences (ViT-B/16, 197 tokens; csrc/vit.cu attention_fwd_rows_kernel,
attention_bwd_q_kernel, attention_bwd_kv_kernel) executed UNCHANGED on host threads by a small CUDA execution-model shim
(tests/emu/cuda_emu.h) and compared with a float64 softmax attention and its gradient.  The kernel text is cut out of vit.cu at
test time, so this checks the index arithmetic and data flow of the code that ships, without a GPU; the whole-sequence kernels
that the GPU tests already verify run through the same shim as a check of the shim itself."""
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


def extract(vit_cu, common_cuh):
    parts = ['namespace smc {', _cut(common_cuh, r'^__device__ __forceinline__ float warp_sum\('),
             _cut(vit_cu, r'^__device__ __forceinline__ void store_split\(')]
    for k in KERNELS:
        body = _cut(vit_cu, r'^__global__ void __launch_bounds__\(256\) ' + k + r'\(')
        assert body.count('extern __shared__ float sm[];') == 1
        parts.append(body.replace('extern __shared__ float sm[];', 'float* sm = emu_smem;'))
    parts.append('}  // namespace smc')
    parts.append(_cut(vit_cu, r'^static int attention_block_rows\('))
    return '\n'.join(parts) + '\n'


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_tiled_attention_kernels_on_the_cpu_shim(tmp_path):
    csrc = os.path.join(ROOT, 'stylemc_b200', 'csrc')
    inc = extract(open(os.path.join(csrc, 'vit.cu')).read(), open(os.path.join(csrc, 'common.cuh')).read())
    (tmp_path / 'kernels_extracted.inc').write_text(inc)
    exe = str(tmp_path / 'attention_emu')
    subprocess.check_call(['g++', '-std=c++20', '-O2', '-pthread', '-Wno-unknown-pragmas', '-I', str(tmp_path), '-I', EMU,
                           os.path.join(EMU, 'attention_main.cpp'), '-o', exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok  ') == 5 and 'FAIL' not in r.stdout
