"""Per-launch ncu counters (tools/round_profile.sh legs 2/3) -> a tracked markdown summary under profiles/.
usage: python tools/counters_md.py <counters.csv> <out.md> <title> [min_ms]"""
import collections, csv, re, sys

src, dst, title = sys.argv[1:4]
min_ms = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
lines = [l for l in open(src) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(int(r['ID']), {'name': re.sub(r'^void |\(.*', '', r['Kernel Name'])})
    v, u = float(r['Metric Value'].replace(',', '')), r['Metric Unit']
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}.get(u, 1.0)
    d[r['Metric Name']] = v * scale
T = 'gpu__time_duration.sum'; R = 'dram__bytes_read.sum'; W = 'dram__bytes_write.sum'
D = 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'; P = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
X = 'l1tex__m_xbar2l1tex_read_bytes.sum'; L = 'lts__throughput.avg.pct_of_peak_sustained_elapsed'
with open(dst, 'w') as f:
    f.write(f'# {title}\n\nCommand: `ncu --profile-from-start off --clock-control none -k regex:... --metrics {T},{R},{W},{D},{P},{X},{L} --csv '
            f'python bench.py --steps 1 --warmup 1 --batch 64 --micro-batch 64 --no-cpu-baseline --profile-step` (tools/round_profile.sh; bench configuration: '
            f'1024 px, 64 seeds, x3p).  Launches of at least {min_ms} ms are listed one by one; the per-kernel totals cover all launches.\n\n')
    f.write('| # | kernel | ms | DRAM read GB | DRAM write GB | DRAM GB/s | DRAM % of peak | tensor pipe active % | L2->SM GB | L2 % |\n|---|---|---|---|---|---|---|---|---|---|\n')
    for i, d in launch.items():
        if d.get(T, 0) >= min_ms:
            gbs = (d.get(R, 0) + d.get(W, 0)) / (d[T] / 1e3)
            f.write(f"| {i} | `{d['name'][:60]}` | {d[T]:.3f} | {d.get(R, 0):.2f} | {d.get(W, 0):.2f} | {gbs:.0f} | {d.get(D, 0):.1f} | {d.get(P, 0):.1f} | {d.get(X, 0):.2f} | {d.get(L, 0):.1f} |\n")
    agg = collections.OrderedDict()
    for d in launch.values():
        a = agg.setdefault(d['name'], [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get(T, 0); a[2] += d.get(R, 0) + d.get(W, 0); a[3] += d.get(P, 0) * d.get(T, 0); a[4] += d.get(D, 0) * d.get(T, 0)
    f.write('\n| kernel (all launches) | launches | total ms | DRAM GB | DRAM GB/s | time-weighted DRAM % | time-weighted tensor pipe % |\n|---|---|---|---|---|---|---|\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'| `{k[:70]}` | {a[0]} | {a[1]:.2f} | {a[2]:.2f} | {a[2] / (a[1] / 1e3):.0f} | {a[4] / a[1]:.1f} | {a[3] / a[1]:.1f} |\n')
    tt = sum(a[1] for a in agg.values())
    f.write(f'\n{len(launch)} launches, {tt:.1f} ms in total; time-weighted tensor pipe activity {sum(a[3] for a in agg.values()) / tt:.1f} %, '
            f'time-weighted DRAM utilisation {sum(a[4] for a in agg.values()) / tt:.1f} % (ncu replays are cold-cache and serialised).\n')
print(open(dst).read()[:2500])
