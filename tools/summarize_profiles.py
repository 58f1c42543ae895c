"""Turn gpurun_out/launches.csv (+ an ncu --set full report) into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> [ncu-rep]"""
import collections, csv, os, re, subprocess, sys

tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else None
out = open(f'profiles/{tag}_launches.md', 'w')
lines = [l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows:
    n = re.sub(r'^void |\(.*', '', r['Kernel Name'])
    tot[n] += us(r); cnt[n] += 1
T = sum(tot.values())
out.write(f'# ncu launch list, one find_direction step (tag {tag})\n\n')
B = os.environ.get('BATCH', '64')
out.write(f'Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 1 --batch {B} '
          f'--micro-batch {B} --no-cpu-baseline --profile-step` (tools/launch_profile.sh; one step of {B} seeds at 1024 px, precision x3p).\n'
          'Per-launch times are cold-cache and serialised: compare SHARES.\n\n')
out.write(f'{len(rows)} launches, {T/1e3:.2f} ms summed device time\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    out.write(f'| `{k[:90]}` | {cnt[k]} | {v:.1f} | {100*v/T:.1f}% |\n')
ig = sum(v for k, v in tot.items() if 'igemm_kernel' in k or 'hconv_kernel' in k)
out.write(f'\nsmc_igemm family (hconv_kernel + igemm_kernel) share of the step: {100*ig/T:.1f}% (bench.py roofline.family.share_of_step must agree).\n')
out.close()
os.system(f'cp gpurun_out/launches.csv profiles/{tag}_launches.csv')
if rep:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]
    keep = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
            'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
    with open(f'profiles/{tag}_top_kernel.md', 'w') as f:
        f.write(f'# ncu --set full, top kernel captures (tag {tag})\n\n| metric | unit | ' + ' | '.join(f'launch {i}' for i in range(len(data))) + ' |\n|---|---|' + '---|' * len(data) + '\n')
        for k in keep:
            for i, h in enumerate(hdr):
                if h == k:
                    f.write(f'| {h} | {units[i]} | ' + ' | '.join(d[i][:70] for d in data) + ' |\n')
print(open(f'profiles/{tag}_launches.md').read()[:3000])
