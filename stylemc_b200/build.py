"""Ahead-of-time build of libstylemc_b200.so (sm_100a) with nvcc; no JIT, no torch headers.

The reference JIT-builds its plugins per process (torch_utils/custom_ops.py:46-124); here the library is
built once in-tree (`python -m stylemc_b200.build`, also run by __graft_entry__.build()) and loaded with
ctypes, so the built .so travels with the repo snapshot and nothing depends on a per-user cache.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, 'csrc')
INCLUDE = os.path.join(os.path.dirname(ROOT), 'include')
LIB = os.path.join(ROOT, 'libstylemc_b200.so')
SOURCES = ['misc.cu', 'bias_act.cu', 'upfirdn2d.cu', 'igemm.cu', 'hconv.cu', 'synth.cu', 'vit.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
              '-I' + INCLUDE, '-I' + CSRC]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', 'nvcc')
    deps = [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'tc.cuh'), os.path.join(INCLUDE, 'stylemc_b200.h')]
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src[:-3] + '.o')
        objs.append(o)
        if force or _newer(s, o) or any(_newer(d, o) for d in deps):
            cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {src}\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if force or procs or not os.path.exists(LIB) or any(_newer(o, LIB) for o in objs):
        subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
