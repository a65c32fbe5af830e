#!/bin/bash
# Round evidence in one GPU call: ncu full capture of the hot kernels, launch list of the bench command,
# and the bench line itself (never a number taken under ncu).
set -x
ncu --set full --clock-control none --import-source on \
    --kernel-name regex:"assemble_tasks|quad_geometry|quad_adjoint|node_gather|bsr_spmv_kernel|cg_update|cg_direction" \
    --launch-skip 6 --launch-count 14 -o gpurun_out/prof_r1d -f python scripts/prof_kernels.py 1024 4 2>&1 | tail -2
ncu -i gpurun_out/prof_r1d.ncu-rep --page raw --csv > gpurun_out/raw_r1d.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/raw_r1d.csv > gpurun_out/r1d_ncu_kernels_1024.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_bench_1024.csv \
    python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r1d_launches_bench_1024.csv > gpurun_out/r1d_launches_bench_1024.txt
cat gpurun_out/r1d_launches_bench_1024.txt
python bench.py --steps 20 > gpurun_out/r1d_bench_n1.json 2> gpurun_out/r1d_bench_n1.err
tail -c 300 gpurun_out/r1d_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r1d_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'clocks', 'gpu_launches')})
print(d['e2e']); print(d['roofline']); print(d['roofline_assembly']); print(d['grad_eval']); print(d.get('cpu_baseline'))
PY
