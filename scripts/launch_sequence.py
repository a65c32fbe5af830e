"""One iteration's kernel sequence (name, grid, microseconds) from an `ncu --metrics gpu__time_duration.sum --csv`
launch list:  python scripts/launch_sequence.py LIST.csv [FIRST] [COUNT]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum':
        continue
    v = float(r[idx['Metric Value']].replace(',', ''))
    v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(r[idx['Metric Unit']], 1e-3)
    seq.append((r[idx['Kernel Name']].split('(')[0][-48:], r[idx['Grid Size']], v))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else len(seq)
print(f'{len(seq)} launches, {sum(v for *_, v in seq):.1f} us in total; launches {first}..{first + count - 1}:')
tot = 0.0
for i, (n, g, v) in enumerate(seq[first:first + count]):
    tot += v
    print(f'{first + i:4d}  {n:50s} {g:16s} {v:9.1f} us')
print(f'      sum of the listed launches {tot:.1f} us')
