"""Opcode histogram (by executed warp instructions) of an `ncu --page source --csv` export. usage: python tools/ncu_sass.py source.csv [top]"""
import csv, sys, collections
rr = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith('==')]
# find header rows (each kernel has 'Kernel Name' row then header row)
i = 0
while i < len(rr):
    if rr[i][0] == 'Kernel Name':
        name = rr[i][1][:60]
        hdr = rr[i + 1]
        si, ei = hdr.index('Source'), hdr.index('Instructions Executed')
        sm = hdr.index('# Samples')
        j = i + 2
        hist, samp = collections.Counter(), collections.Counter()
        tot = 0
        while j < len(rr) and rr[j][0] != 'Kernel Name':
            r = rr[j]
            try:
                n = int(r[ei]); s = int(r[sm])
            except ValueError:
                j += 1; continue
            toks = r[si].split()
            op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
            op = '.'.join(op.split('.')[:3])
            hist[op] += n; samp[op] += s; tot += n
            j += 1
        print(f'== {name}: {tot} warp instructions, {j - i - 2} SASS lines')
        for op, n in hist.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
            print(f'  {op:28s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]}')
        i = j
    else:
        i += 1
