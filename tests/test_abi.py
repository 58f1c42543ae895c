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
