"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[idx['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[idx['Kernel Name']].split('(')[0][-60:]
    v = float(r[idx['Metric Value']].replace(',', ''))
    unit = r[idx['Metric Unit']]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}.get(unit, 1e-6)
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
T = sum(v for _, v in tot.values())
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{v:10.3f} ms {100 * v / T:5.1f}%  x{n:5d}  {k}')
print(f'{T:10.3f} ms total')
