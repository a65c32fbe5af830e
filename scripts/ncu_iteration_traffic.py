"""DRAM traffic and time of ONE multigrid-PCG iteration from an `ncu --set full --page raw --csv` dump of
scripts/mg_profile.py (profiled range = `iters` iterations of jsso_pcg):  python scripts/ncu_iteration_traffic.py RAW.csv ITERS"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
iters = int(sys.argv[2])
hdr = rows[0]; units = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[idx[k]].replace(',', '')); u = units[idx[k]]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1.0)
per = {}
tot_b = tot_t = 0.0
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0][-40:]
    b = val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
    t = val(r, 'gpu__time_duration.sum')
    a = per.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += b; a[2] += t
    tot_b += b; tot_t += t
out = {'iterations_profiled': iters, 'dram_bytes_per_iteration': tot_b / iters, 'kernel_us_per_iteration_under_ncu': tot_t / iters,
       'by_kernel_per_iteration': {k: {'launches': v[0] / iters, 'dram_bytes': v[1] / iters, 'us': v[2] / iters} for k, v in per.items()}}
print(json.dumps(out, indent=1))
