#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library: counts of the tcgen05 / TMEM / TMA / mbarrier instructions in every kernel of
libstylemc_b200.so (cuobjdump -sass), written as a markdown table.  No GPU needed.

    python tools/sass_table.py > profiles/<tag>_sass_hist.md

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (fp16 kind), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,
UTMALDG = TMA tensor load (cp.async.bulk.tensor), SYNCS = mbarrier ops, FFMA2 = packed fp32x2 FMA (sm_100)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'stylemc_b200', 'libstylemc_b200.so')
COLS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'SYNCS', 'HMMA', 'FFMA2', 'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'ATOMG', 'RED', 'SHFL']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            op = m.group(1)
            cur[op] += 1
            cur['_total'] += 1
    names = demangle(list(kernels))
    rows = []
    for k, c in kernels.items():
        n = re.sub(r'^void ', '', names.get(k, k))
        n = re.sub(r'\(.*$', '', n).replace('smc::', '')
        rows.append((n, c))
    rows.sort(key=lambda r: (-r[1]['UTCHMMA'], -r[1]['UTMALDG'], r[0]))
    print('# SASS instruction counts per kernel, libstylemc_b200.so (sm_100a)\n')
    print('`python tools/sass_table.py` (cuobjdump -sass of the built library; static counts, loops not unrolled count once).')
    print('UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, SYNCS = mbarrier.\n')
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    print(f'{len(rows)} kernels, {tot["_total"]} instructions; library totals: ' + ', '.join(f'{k} {tot[k]}' for k in COLS if tot[k]) + '\n')
    print('| kernel | instr | ' + ' | '.join(COLS) + ' |')
    print('|---|---|' + '---|' * len(COLS))
    for n, c in rows:
        print(f'| `{n}` | {c["_total"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in COLS) + ' |')


if __name__ == '__main__':
    sys.exit(main())
