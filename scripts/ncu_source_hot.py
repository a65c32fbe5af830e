"""Aggregate an `ncu --page source --csv` dump (SASS view) into stall-reason totals and
the hottest instruction ranges.  usage: ncu_source_hot.py file.csv [n_top]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith('stall_')]
tot = {c: 0 for c in stall_cols}
samples = 0
inst = 0
per = []
for r in data:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx['# Samples']])
        ie = int(r[idx['Instructions Executed']])
    except ValueError:
        continue
    samples += s
    inst += ie
    for c in stall_cols:
        try:
            tot[c] += int(r[idx[c]])
        except ValueError:
            pass
    per.append((s, ie, r[idx['Source']][:90]))
print('samples', samples, 'warp instructions', inst)
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f'  {c:28s} {v:8d} {100.0 * v / max(samples, 1):5.1f}%')
# opcode mix
mix = {}
for s, ie, src in per:
    op = src.split()[0] if src.split() else '?'
    if op.startswith('@'):
        op = src.split()[1] if len(src.split()) > 1 else op
    op = op.split('.')[0]
    mix[op] = mix.get(op, 0) + ie
print('opcode mix (warp instructions):')
for op, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]:
    print(f'  {op:10s} {v:10d} {100.0 * v / max(inst, 1):5.1f}%')
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print('hottest instructions by samples:')
for s, ie, src in sorted(per, key=lambda t: -t[0])[:n]:
    print(f'  {s:7d} {ie:9d}  {src}')
