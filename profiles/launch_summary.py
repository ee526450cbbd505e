"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: per-kernel totals of the last
complete render step and the per-lane (stream) sequence of cull launch times.  Per-launch times under ncu are
serialised and cold-cache: compare SHARES, not absolutes.  usage: launch_summary.py launches.csv"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); si = hdr.index('Stream')
seq = [(r[ki].split('(')[0].replace('void rt::', '').replace('rt::', '').split('<')[0], float(r[vi].replace(',', '')) / 1e3, r[si]) for r in rows[1:]]
inits = [i for i, (k, _, _) in enumerate(seq) if k == 'wf_init']
print(f"{len(seq)} launches, {len(inits)} renders in the list; last complete render step:")
step = seq[inits[-2]:inits[-1]]
tot = defaultdict(float); cnt = defaultdict(int)
for k, us, _ in step: tot[k] += us; cnt[k] += 1
total = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"  {k:44s} {cnt[k]:4d} launches {tot[k]:9.0f} us  {100 * tot[k] / total:5.1f} %")
print(f"  {'sum (serialised under ncu)':44s} {len(step):4d} launches {total:9.0f} us")
by = defaultdict(list)
for k, us, s in step:
    if k == 'wf_cull': by[s].append(round(us))
for s, v in by.items(): print(f"  lane on stream {s}: wf_cull us per iteration {v}")
