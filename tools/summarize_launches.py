"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
fam = collections.defaultdict(list)
rows = []
for row in r:
    if len(row) <= vi:
        continue
    name = row[ki]
    v = float(row[vi].replace(',', '')) / 1000
    short = name.split('(')[0].split('::')[-1]
    if 'igemm_kernel' in name:
        short = 'igemm_' + name.split('<')[1].split('>')[0].replace('(int)', '')
    fam[short].append(v)
    rows.append((v, short, row[gi]))
tot = sum(sum(v) for v in fam.values())
print('total us', round(tot, 1), 'kernels', len(rows))
for k, v in sorted(fam.items(), key=lambda kv: -sum(kv[1])):
    print(f'{k:34s} n={len(v):4d} sum={sum(v):8.1f} share={100 * sum(v) / tot:5.1f}% avg={sum(v) / len(v):6.1f} '
          f'max={max(v):6.1f} min={min(v):6.1f}')
ig = [x for x in rows if x[1].startswith('igemm')]
print('slowest igemm (us, grid):', [(round(x[0], 1), x[2]) for x in sorted(ig, key=lambda t: -t[0])[:16]])
