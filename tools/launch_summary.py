"""Per-kernel summary of the LAST complete step of an ncu launch list (gpu__time_duration.sum CSV); steps are delimited by
k_sys_ptr (first kernel of the neighbour build).    python tools/launch_summary.py X.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iu = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit'))
L = []
for row in r:
    v = float(row[iv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row[iu], 1.0)
    L.append((row[ik], v))
st = [i for i, (n, v) in enumerate(L) if 'k_sys_ptr' in n]
step = L[st[-2]:st[-1]] if len(st) >= 2 else L
tot = collections.OrderedDict()
for n, v in step:
    n = n.split('(')[0].replace('void ', '').replace('<unnamed>::', '').replace('at::native::', '')[:76]
    t = tot.setdefault(n, [0, 0.0]); t[0] += 1; t[1] += v
S = sum(t[1] for t in tot.values())
print(f'# {len(step)} launches per step, {S:.0f} us summed kernel time (ncu: serialised, cold caches)')
print(f'{"kernel":78s} {"n":>4s} {"us":>9s} {"share":>6s}')
for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{n:78s} {t[0]:4d} {t[1]:9.1f} {100 * t[1] / S:5.1f}%')
