#!/bin/bash
# Round 2, GPU call 18 (one GPU): final single-GPU evidence with the code as committed -- full GPU test suite, smoke,
# numeric setup by phase, launch list of the M1 bench command, the bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2zz_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2zz_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
JSSO_MG_TIMING=1 timeout 600 python scripts/mg_profile.py 1024 3 1 setup > gpurun_out/r2zz_setup_phases.txt 2>&1; echo "phases rc=$?"; grep JSSO_MG_TIMING gpurun_out/r2zz_setup_phases.txt | tail -2 | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zz_launches_bench_1024.csv \
    python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2zz_bench_under_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2zz_launches_bench_1024.csv > gpurun_out/r2zz_launches_bench_1024.txt; head -8 gpurun_out/r2zz_launches_bench_1024.txt
python bench.py > gpurun_out/r2zz_bench_n1.json 2> gpurun_out/r2zz_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2zz_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2zz_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'gpu_launches', 'clocks')})
print(d['e2e']['ms_per_step']); print(d['m2']); print(d['roofline_pcg_iteration']['frac'], d['roofline']['frac'], d['roofline_assembly']['frac'], d['roofline_adjoint']['frac'], d['roofline_spmv']['frac'])
print(d['grad_eval']['stage_s'], d['grad_eval']['seconds_each']); print(d.get('batch_eval', {}).get('designs_per_s')); print(d.get('topo_eval', {}).get('s_per_iteration'))
PY
