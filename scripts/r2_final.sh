#!/bin/bash
# Round 2, final single-GPU evidence: full GPU test suite, smoke, compute-sanitizer, ncu launch lists and full
# captures of the final kernels, the bench line (with cpu_baseline) and the reference arm.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2z_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2z_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py > gpurun_out/r2z_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2z_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_run.py > gpurun_out/r2z_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2z_sanitizer_racecheck.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2z_launches_mg.csv python scripts/mg_profile.py 1024 3 > gpurun_out/r2z_mgprof.log 2>&1; echo "launch list rc=$?"
python scripts/launch_sequence.py gpurun_out/r2z_launches_mg.csv 4 24 > gpurun_out/r2z_launches_mg_iteration.txt; tail -3 gpurun_out/r2z_launches_mg_iteration.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2z_mg -f \
   python scripts/mg_profile.py 1024 2 > gpurun_out/r2z_mgprof_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2z_mg.ncu-rep --page raw --csv > gpurun_out/r2z_mg_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2z_mg_raw.csv > gpurun_out/r2z_ncu_mg_1024.txt 2>&1
python scripts/ncu_iteration_traffic.py gpurun_out/r2z_mg_raw.csv 2 > gpurun_out/r2z_ncu_iteration_traffic.json 2>&1
rm -f gpurun_out/r2z_mg.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_bench_1024.csv \
    python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2z_bench_under_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2z_launches_bench_1024.csv > gpurun_out/r2z_launches_bench_1024.txt; cat gpurun_out/r2z_launches_bench_1024.txt | head -8
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2z_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'gpu_launches', 'clocks')})
print(d['e2e']); print(d['m2']); print(d['roofline_pcg_iteration']['frac'], d['roofline']['frac'], d['roofline_assembly']['frac'], d['roofline_adjoint']['frac'], d['roofline_spmv']['frac'])
print(d['grad_eval']['stage_s'], d['grad_eval']['seconds_each']); print(d.get('batch_eval')); print(d.get('topo_eval')); print(d.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "reference arm rc=$?"; cut -c1-600 gpurun_out/r2z_bench_reference.json
