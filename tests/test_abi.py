"""CPU: the C-ABI library loads and exports exactly the entry points include/stylemc_b200.h declares, and the ctypes mirror
(stylemc_b200/_lib.py) agrees with the header on every argument list.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'stylemc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    out = {}
    for m in re.finditer(r'\bint\s+(smc_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace('\n', ' ').split(',') if a.strip() and a.strip() != 'void']
        out[m.group(1)] = args
    return out


def kind(arg):
    if '*' in arg:
        return 'p'
    if arg.startswith('float'):
        return 'f'
    if arg.startswith('int64_t'):
        return 'q'
    return 'i'


def test_library_exports_every_declared_symbol():
    from stylemc_b200 import _lib, build
    build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    decl = header_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    assert handle.smc_abi_version() == 1


def test_ctypes_signatures_match_header():
    from stylemc_b200 import _lib
    for name, args in header_functions().items():
        want = ''.join(kind(a) for a in args)
        assert _lib.SIGNATURES[name].replace(' ', '') == want, (name, _lib.SIGNATURES[name], want)


def test_struct_layouts_match_header():
    from stylemc_b200 import _lib
    # smc_igemm_tap: 4 x int32; descriptor holds 32 taps; sizes are what nvcc lays out for the same field lists
    assert ctypes.sizeof(_lib.Tap) == 16
    assert _lib.IgemmDesc.taps.size == 32 * 16
    assert ctypes.sizeof(_lib.UpfirdnParams) % 8 == 0
    src = open(os.path.join(ROOT, 'include', 'stylemc_b200.h')).read()
    assert f'#define SMC_IGEMM_MAX_TAPS {_lib.MAX_TAPS}' in src


def test_struct_offsets_match_a_c_compiler(tmp_path):
    """Every field of the three structs that cross the boundary sits where gcc puts it for include/stylemc_b200.h (sizeof + offsetof
    printed by a tiny C program); catches a ctypes mirror that drifts from the header."""
    import shutil
    import subprocess
    from stylemc_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('no C compiler')
    structs = {'smc_igemm_epilogue': _lib.Epilogue, 'smc_igemm_desc': _lib.IgemmDesc, 'smc_upfirdn2d_params': _lib.UpfirdnParams,
               'smc_igemm_tap': _lib.Tap}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "stylemc_b200.h"', 'int main(void) {']
    for cname, st in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == ctypes.sizeof(st), (cname, got[cname], ctypes.sizeof(st))
        for fname, _ in st._fields_:
            assert int(got[f'{cname}.{fname}']) == getattr(st, fname).offset, (cname, fname)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'stylemc_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f'{fn} imports the oracle'


def test_cpu_tensors_are_refused():
    import torch
    from stylemc_b200.ops import bias_act, upfirdn2d
    with pytest.raises(RuntimeError):
        bias_act.bias_act(torch.zeros(1, 2, 3, 3), torch.zeros(2), act='lrelu')
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(torch.zeros(1, 2, 4, 4), upfirdn2d.setup_filter([1, 3, 3, 1]))


def test_act_bwd_refuses_channel_counts_whose_lane_groups_do_not_tile_a_warp():
    """C / 8 = 12 lanes per pixel leave 8 lanes of a warp over; they would redo a neighbour's pixel and the style-gradient reductions would
    count it twice (found with tests/test_kernels_emu.py).  The C layer refuses such shapes before any launch: no device needed."""
    from stylemc_b200 import _lib
    fake = 0x1000                                        # never dereferenced on the host
    args = lambda c: (fake, fake, 1, 8, 8, c, fake, 1, fake, 512, None, None, None, 0, 1.0, None, -1.0, fake, fake, fake, fake, 0.2, 1.4, 256.0,
                      fake, fake, fake, fake, None)
    assert _lib.lib().smc_act_bwd(*args(96)) == -2       # SMC_EUNSUPPORTED
    assert _lib.lib().smc_act_bwd(*args(24)) == -1       # below 32 channels: invalid argument, as before
